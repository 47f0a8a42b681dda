"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum): usage launch_times.py launches.csv [skip_first_n_solve_calls]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
seq = [(int(r[ii]), r[ki].split("(")[0], float(r[vi].replace(",", ""))) for r in rows[1:] if r[vi].replace(",", "").replace(".", "").isdigit()]
print("last call, in launch order:")
# the last call = the trailing launches after the last k_sched_hist's preceding solve launch
last = max(i for i, (_, k, _) in enumerate(seq) if k == "k_sched_hist")
start = last - 1
tot = sum(v for _, _, v in seq[start:])
for _, k, v in seq[start:]:
    print("  %-24s %10.3f ms %5.1f%%" % (k, v * 1e-6, 100 * v / tot))
print("  total %.3f ms" % (tot * 1e-6))
