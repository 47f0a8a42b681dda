#!/usr/bin/env python
"""profiles/r2_traffic.json from an ncu launch list of one 2^20 step (tools/launch_split.sh), stamped with the hash of the
kernel sources it was taken with: python tools/make_traffic.py gpurun_out/<tag>_launches.csv out.json"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_times.py"), sys.argv[1]], capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
rec = {"source_hash": bench.source_hash(), "log2n": 20, "kernel": 0, "stop_rule": 0, "dram_bytes_per_step": d["dram_bytes_per_step"],
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the launches of one 2^20 step (tools/launch_split.sh; "
              "profiles/r2_launches.csv): launch A writes 8.5 GB (1 KB of state per model, 4.9-10.5 KB of capture per cacheable "
              "model, the results), the engines read it back",
       "serialised_ms": d["serialised_ms"]}
json.dump(rec, open(sys.argv[2], "w"), indent=1)
print(rec)
