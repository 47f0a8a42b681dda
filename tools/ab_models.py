"""Debug: per-model comparison of kernel=0 (cached) against kernel=2 (full elimination every iteration)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import MOLFILE, draw_params
from radex_emcee_b200 import _lib
from test_gpu_solve import gpu_solve

ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = draw_params(np.random.default_rng(1012), n, 10.926)
kw = {}
if len(sys.argv) > 2:
    kw["maxiter"] = int(sys.argv[2])
a = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, **kw)
print("cache stats", ctx.cache_stats())
b = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, kernel=2, **kw)
with np.errstate(all="ignore"):
    ex = np.nanmax(np.where(b["xpop"] > 1e-12, np.abs(a["xpop"] - b["xpop"]) / b["xpop"], 0), axis=1)
top = np.array([(np.nonzero(np.abs(t) / 2 >= 0.01)[0].max() + 1) if (np.abs(t) / 2 >= 0.01).any() else -1 for t in b["tau"]])
bad = ex > 1e-6
print("models %d, differing %d" % (n, bad.sum()))
print("topthick histogram of differing:", np.bincount(top[bad] + 1, minlength=42))
print("topthick histogram of all      :", np.bincount(top + 1, minlength=42))
for i in np.nonzero(bad)[0][:12]:
    print(i, "T=%.3g n=%.3g N=%.3g" % tuple(P[i]), "niter", a["niter"][i], b["niter"][i], "top", top[i], "err %.2e" % ex[i],
          "x0..3", a["xpop"][i][:4], b["xpop"][i][:4])
