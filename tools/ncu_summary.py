"""Markdown table of one .ncu-rep (`ncu --set full` capture): per kernel launch the duration, DRAM bytes, registers,
pipe utilisations, resident warps, executed warp instructions and issue utilisation.
   python tools/ncu_summary.py <report.ncu-rep> [title]   (appended by hand to profiles/ncu_r1_summary.md)"""
import csv
import subprocess
import sys

COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time [us]"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instrs"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = [(h.index(c), t) for c, t in COLS if c in h]
if len(sys.argv) > 2:
    print("## " + sys.argv[2] + "\n")
print("| " + " | ".join(t + (" [" + units[i] + "]" if units[i] and t in ("dram read", "dram write") else "") for i, t in ix) + " |")
print("|" + "---|" * len(ix))
for r in rows[2:]:
    cells = []
    for i, t in ix:
        v = r[i]
        if t == "kernel":
            v = v.replace("void ", "").split("(")[0][:48]
        else:
            try:
                f = float(v)
                v = ("%d" % f) if f == int(f) and abs(f) < 1e15 else ("%.3g" % f if abs(f) < 100 else "%.1f" % f)
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
