"""Per-kernel totals of an ncu launch list: usage launch_times.py launches.csv
Metrics read: gpu__time_duration.sum (ns) and, when present, dram__bytes_read.sum / dram__bytes_write.sum.
Prints the launches of the LAST solve call (from the launch A that precedes the last k_sched_hist on), each with its
share of the step, and the DRAM bytes of that step; writes nothing."""
import csv
import json
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
launch = OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "nsecond": 1.0, "ns": 1.0}.get(unit, 1.0)
    d = launch.setdefault(int(r[ii]), {"kernel": r[ki].split("(")[0]})
    d[r[mi]] = v * scale
seq = list(launch.values())
last = max(i for i, d in enumerate(seq) if d["kernel"] == "k_sched_hist")
step = [d for d in seq[last - 1:] if d["kernel"] != "k_fp64_peak"]
tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in step)
dram = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in step)
print("last call, in launch order:")
for d in step:
    t = d.get("gpu__time_duration.sum", 0.0)
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    print("  %-24s %10.3f ms %5.1f%%  dram %8.3f GB" % (d["kernel"], t * 1e-6, 100 * t / tot, b * 1e-9))
print("  total %.3f ms (serialised, cold cache), dram %.3f GB" % (tot * 1e-6, dram * 1e-9))
print(json.dumps({"dram_bytes_per_step": dram, "serialised_ms": tot * 1e-6}))
