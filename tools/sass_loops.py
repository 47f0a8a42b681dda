#!/usr/bin/env python
"""Backward branches (loops) of one kernel in the in-tree library: python tools/sass_loops.py <kernel-substring> [minlen]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if sys.argv[1] not in f.split("\n", 1)[0]:
        continue
    n = 0
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,8})\*/\s+(.*?);", line)
        if not m:
            continue
        n += 1
        a, t = int(m.group(1), 16), m.group(2)
        if "BRA" in t:
            mm = re.search(r"0x([0-9a-f]+)", t)
            if mm and int(mm.group(1), 16) < a and (a - int(mm.group(1), 16)) // 16 >= minlen:
                print("%#x -> %#x  %d instructions  %s" % (a, int(mm.group(1), 16), (a - int(mm.group(1), 16)) // 16, t[:40]))
    print("total", n)
