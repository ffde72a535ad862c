"""Per-source-line profile of one kernel from an ncu capture taken with --import-source on (kernel built with -lineinfo):
joins the SASS rows of `ncu --page source --csv` with the line table of `nvdisasm -g` on the kernel's cubin and prints
the hottest source lines (executed warp instructions, stall samples).
   python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
import os
ksel = ["-k", os.environ["NCU_K"]] if os.environ.get("NCU_K") else []   # NCU_K=regex:<name> picks one kernel of the report
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
end = next((i for i in range(hdr + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))  # first kernel only
sass = [r for r in rows[hdr + 1:end] if len(r) == len(h)]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = pat in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
ie, ss, si = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for ln, r in zip(lines, sass):
    agg[ln][0] += int(r[ie]); agg[ln][1] += int(r[ss])
    tot_i += int(r[ie]); tot_s += int(r[ss])
src = {}
def text(ln):
    if ln is None:
        return ""
    f, n = ln
    if f not in src:
        import glob
        c = glob.glob("/root/repo/geoa3_b200/csrc/" + f) + glob.glob("/root/repo/include/" + f)
        src[f] = open(c[0]).read().splitlines() if c else []
    return src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
print("total warp instrs %d, samples %d" % (tot_i, tot_s))
for ln, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%s  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), ln[0] if ln else "?", ln[1] if ln else "?", text(ln)))
if len(sys.argv) > 5:  # region summary: "name:lo-hi,name:lo-hi,..." over the kernel's own file; other files listed by name
    regs = [(a.split(":")[0], int(a.split(":")[1].split("-")[0]), int(a.split(":")[1].split("-")[1])) for a in sys.argv[5].split(",")]
    main = max(set(l[0] for l in lines if l), key=lambda f: sum(1 for l in lines if l and l[0] == f))
    ragg = defaultdict(lambda: [0, 0])
    for ln, (i, s) in agg.items():
        name = "?"
        if ln and ln[0] == main:
            name = next((r[0] for r in regs if r[1] <= ln[1] <= r[2]), "other")
        elif ln:
            name = ln[0] + ":" + str(ln[1])
        ragg[name][0] += i; ragg[name][1] += s
    print("---- regions")
    for name, (i, s) in sorted(ragg.items(), key=lambda kv: -kv[1][0]):
        print("%5.1f%% ins %5.1f%% smp  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), name))
