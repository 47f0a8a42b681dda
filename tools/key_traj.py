"""Lead-block key (lvg_v2.cuh: want) after 1, 2, 3, 5, 10, 20 calls of matrix() and at the end, on the GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MOLFILE  # noqa: E402
from radex_emcee_b200 import _lib  # noqa: E402
from test_gpu_solve import gpu_solve  # noqa: E402
from bench import draw, TBG  # noqa: E402

n = 20000
tk, nh2, cd = draw(n, 0)
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)


def key(tau):
    thick = ~(np.abs(tau * 0.5) < np.float32(0.01))
    top = np.where(thick.any(axis=1), thick.shape[1] - np.argmax(thick[:, ::-1], axis=1), -1)
    return np.maximum(3, (top + 1 + 4) >> 2)


keys = {}
for mi in (1, 2, 3, 5, 10, 20, 200):
    out = gpu_solve(ctx, tk, nh2, cd, TBG, maxiter=mi, kernel=3)
    keys[mi] = key(out["tau"])
    niter = out["niter"]
k1, kf = keys[1], keys[200]
print("rows: key after call 0 (3..11); cols: final key (3..11)")
for a in range(3, 12):
    print(a, [int(((k1 == a) & (kf == b)).sum()) for b in range(3, 12)])
for mi in keys:
    print("after", mi, "calls: mean key %.3f  frac<=4 %.3f  equal to final %.3f  below final %.3f above final %.3f"
          % (keys[mi].mean(), (keys[mi] <= 4).mean(), (keys[mi] == kf).mean(), (keys[mi] < kf).mean(), (keys[mi] > kf).mean()))
big_then_small = (k1 > 4) & (kf <= 4)
print("key>4 after call 0 but <=4 at the end: %.3f of models, mean niter %.1f" % (big_then_small.mean(), niter[big_then_small].mean()))
for mi in (2, 3, 5, 10, 20):
    print("  of those, already <=4 after %d calls: %.3f" % (mi, (keys[mi][big_then_small] <= 4).mean()))
