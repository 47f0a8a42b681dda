"""Dev tool (GPU box): do two builds of libradex_b200 give the same bits?  usage: lib_equal.py libA.so libB.so [n]
Each library solves the same config-2 draws (all three geometries, scheduled and single launch) in its own process."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MOLFILE, draw_params
from radex_emcee_b200 import _lib
from test_gpu_solve import gpu_solve
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
n = int(sys.argv[2]); out = {}
for method, tbg, kernel in ((2, 10.926, 0), (2, 2.7315, 3), (1, 2.7315, 0), (3, 10.926, 0)):
    P = draw_params(np.random.default_rng(7 + method), n, tbg)
    got = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], tbg, method, kernel=kernel)
    for k in got: out["%%s_%%d_%%d" %% (k, method, kernel)] = got[k]
np.savez(sys.argv[1], **out)
'''
libs, n = sys.argv[1:3], (sys.argv[3] if len(sys.argv) > 3 else "20000")
target = os.path.join(ROOT, "radex_emcee_b200", "libradex_b200.so")
keep = open(target, "rb").read()
res = []
try:
    for i, lib in enumerate(libs):
        open(target, "wb").write(open(lib, "rb").read())
        f = "/tmp/lib_equal_%d.npz" % i
        subprocess.check_call([sys.executable, "-c", CHILD % ROOT, f, n])
        res.append(np.load(f))
finally:
    open(target, "wb").write(keep)
bad = [k for k in res[0].files if not np.array_equal(res[0][k], res[1][k], equal_nan=True)]
print("arrays compared:", len(res[0].files), "differing:", bad if bad else "none")
for k in bad:
    a, b = res[0][k], res[1][k]
    with np.errstate(all="ignore"):
        print(" ", k, "max rel diff", np.nanmax(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)), "entries", int((a != b).sum()))
