"""Static SASS opcode counts per kernel of libgeoa3_b200.so (cuobjdump -sass), as a markdown table: evidence for the
packed-fp32 / TMA bulk copy / warp-reduction opcodes and for the ABSENCE of float atomics and tensor-core opcodes.
    python tools/sass_table.py > profiles/sass_<round>.md"""
import collections
import os.path as osp
import re
import subprocess
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else osp.join(ROOT, "geoa3_b200", "libgeoa3_b200.so")
OPS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX", "REDUX", "CREDUX", "SHF", "LDS", "STS", "LDG", "STG", "UBLKCP", "SYNCS", "ATOMS",
       "ATOMG", "RED", "MATCH", "VOTE", "SHFL", "BAR", "HMMA", "UTCMMA"]
elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout.strip()
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("# SASS evidence (`cuobjdump -sass geoa3_b200/libgeoa3_b200.so`, built by `python -m geoa3_b200.build`; table by tools/sass_table.py)\n")
print("Every cubin in the library is `sm_100a` (`cuobjdump -lelf`):\n\n```\n" + elf + "\n```\n")
print("Opcode counts per kernel (static SASS instructions). `UBLKCP` = the TMA bulk copy (`cp.async.bulk`), `SYNCS` = its "
      "mbarrier, `REDUX` = warp reductions, `FFMA2/FADD2/FMUL2` = packed fp32; `ATOMS` are INTEGER shared atomics (slot "
      "hand-out of the counting sorts); there is no float atomic (`ATOMS.*F32`, `RED.*F32`, `ATOMG.*F32`: see the last "
      "line) and no tensor-core opcode (`HMMA`/`UTC*MMA`) by design: the 3-wide contraction is not a GEMM (north star).\n")
print("| kernel | total | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
cur, cnt, tot, fatom = None, None, 0, 0
rows = []


def flush():
    if cur:
        rows.append((cur, tot, cnt))


for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("(bool)", "").replace("(int)", "")
        cnt, tot = collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        tot += 1
        base = op.split(".")[0]
        if base in OPS:
            cnt[base] += 1
        if base in ("ATOMS", "ATOMG", "RED", "ATOM") and (".F32" in op or ".F16" in op or ".F64" in op):
            fatom += 1
flush()
for name, t, c in rows:
    print("| `%s` | %d | " % (name, t) + " | ".join(str(c.get(o, 0)) for o in OPS) + " |")
print("\nFloat atomics (any `ATOM*`/`RED` with an `.F16/.F32/.F64` type) in the whole library: **%d**." % fatom)
