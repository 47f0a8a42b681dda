#!/usr/bin/env python
"""Where does ptxas spill?  usage: tools/spills.py [kernel-substring]  (reads the in-tree libradex_b200.so)
Prints, per source line, the number of local-memory stores/loads (STL/LDL) and the instruction count of the kernel."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else "k_lvg_solve_v2"
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin") and "moldata" not in f][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, infn, n = None, False, 0
st, ld = collections.Counter(), collections.Counter()
ops = collections.Counter()
for line in sass.splitlines():
    if line.startswith(".text."):
        infn = want in line
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        n += 1
        op = m.group(1)
        ops[op.split(".")[0]] += 1
        if op.startswith("STL"):
            st[cur] += 1
        if op.startswith("LDL"):
            ld[cur] += 1
print("kernel *%s*: %d SASS instructions" % (want, n))
print("top opcodes:", ops.most_common(14))
for name, c in (("STL", st), ("LDL", ld)):
    print(name, sum(c.values()), "->", sorted(c.items(), key=lambda kv: (kv[0] or ("", 0))))
