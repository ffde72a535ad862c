"""Prints the hottest SASS lines (warp-stall samples) of one kernel from an .ncu-rep:
   python tools/ncu_hot.py <report.ncu-rep> <kernel-regex> [top=25] [launch-skip=0]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
si, ni, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
body = [r for r in rows[hdr_i + 1:] if len(r) > ni and r[0].startswith("0x")]
tot = sum(int(r[ni] or 0) for r in body)
print(rows[0][1][:100], "| total samples", tot, "| SASS lines", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ni] or 0))[:top]
for i in sorted(order):
    r = body[i]
    print("%5d %6.2f%% exec=%-8s %s" % (i, 100.0 * int(r[ni] or 0) / max(tot, 1), r[ei], r[si].strip()[:100]))
