#!/usr/bin/env python
"""Static SASS count per source line and opcode class of one kernel (no GPU needed):
  python tools/sass_lines.py <kernel-substring> [lo_hex hi_hex] [topN]
nvdisasm --print-line-info on the cubin of the in-tree library; an address window limits the count to a loop."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")
kern = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
topn = int(sys.argv[-1]) if len(sys.argv) in (3, 5) else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin") and "moldata" not in f][0]
lines = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
CLS = {"FP64": ("DFMA", "DADD", "DMUL", "DSETP", "DMMA"), "MUFU": ("MUFU",), "UMOV": ("UMOV",), "MOV": ("MOV", "IMAD.MOV", "CS2R"),
       "LDS/STS": ("LDS", "STS"), "SHFL": ("SHFL",), "CTRL": ("BRA", "BSSY", "BSYNC", "WARPSYNC", "NOP", "CALL", "RET", "BREAK", "EXIT")}
def cls(op):
    for k, v in CLS.items():
        if any(op.startswith(p) for p in v):
            return k
    return "INT/SEL"
cur = ("?", 0)
per = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for l in lines[start + 1:]:
    if l.startswith(".text.") and kern not in l:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m and lo <= int(m.group(1), 16) < hi:
        c = cls(m.group(2))
        per[cur][c] += 1
        tot[c] += 1
n = sum(tot.values())
print("instructions", n, dict(tot))
src = {}
for (f, ln), c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:topn]:
    if f not in src:
        p = [os.path.join(ROOT, "radex_emcee_b200", "csrc", f)]
        src[f] = open(p[0]).read().split("\n") if os.path.exists(p[0]) else []
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ""
    print("%5d %5.1f%%  %s:%d  %s  | %s" % (sum(c.values()), 100.0 * sum(c.values()) / n, f, ln, dict(c), text))
