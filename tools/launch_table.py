"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one attack step captured with
--profile-from-start off + tools/profile_step.py).   python tools/launch_table.py <launches.csv> [rows]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) == len(h)]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, t in data:
    own = "geoa3::" in name
    short = name.replace("void ", "").split("(")[0][:64]
    agg[("own  " if own else "torch") + " " + short][0] += 1
    agg[("own  " if own else "torch") + " " + short][1] += t
tot = sum(v[1] for v in agg.values())
own_t = sum(v[1] for k, v in agg.items() if k.startswith("own"))
print("%d launches, %.1f us in total (cold-cache, serialised); own kernels %.1f us = %.1f %%" % (len(data), tot / 1e3, own_t / 1e3, 100 * own_t / tot))
print("| us | share | launches | kernel |\n|---|---|---|---|")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("| %.1f | %.1f %% | %d | `%s` |" % (t / 1e3, 100 * t / tot, c, name))
