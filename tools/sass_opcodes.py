#!/usr/bin/env python
"""SASS opcode counts per kernel of the built library (cuobjdump -sass), as a markdown table:
  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md
What to look for: DFMA/DMUL/DADD (the FP64 vector pipe the solve lives on), DMMA (FP64 tensor-core MMAs of the rank-4
panel updates), UBLKCP + SYNCS (TMA bulk copy + mbarrier: the collisional matrix restored from L2), SHFL (pivot
broadcasts, reductions), LDS/STS (rate matrix in shared memory); no HMMA/UTC*MMA (there is no FP64 tcgen05 kind)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "UBLKCP", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMG", "BAR"]
rows, tot = [], collections.Counter()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    rows.append((name, sum(ops.values()), ops))
    tot.update(ops)
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.split("\n")
print("# SASS opcode counts, libradex_b200.so (sm_100a), `tools/sass_opcodes.py`\n")
print(__doc__.split("What to look for:")[1].strip().replace("\n", " ") + "\n")
print("| kernel | instructions | " + " | ".join(want) + " |")
print("|---|---|" + "---|" * len(want))
for (name, n, ops), d in sorted(zip(rows, dem), key=lambda r: -r[0][1]):
    short = re.sub(r"\(.*", "", d).replace("void ", "")
    print("| `%s` | %d | " % (short, n) + " | ".join(str(ops.get(w, 0)) for w in want) + " |")
print("| **all** | %d | " % sum(r[1] for r in rows) + " | ".join(str(tot.get(w, 0)) for w in want) + " |")
other = [k for k in tot if re.match(r"(HMMA|UTC|WGMMA|HGMMA|IMMA)", k)]
print("\nTensor-core opcodes other than DMMA: %s" % (", ".join(other) if other else "none"))
