"""Small scheduled solve + pipelined lnprob for compute-sanitizer (memcheck / racecheck): tools/sanitize.py [n] [maxiter]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MOLFILE, draw_params  # noqa: E402
from radex_emcee_b200 import _lib  # noqa: E402
from test_gpu_solve import gpu_solve  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
P = draw_params(np.random.default_rng(5), n, 10.926)
os.environ["RB_PARK_MAX"] = "7"   # all five cached-engine kernels
a = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, maxiter=maxiter)
b = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, maxiter=maxiter, kernel=3)
for k in a:
    assert np.array_equal(a[k], b[k], equal_nan=True), k
print("ok", n, maxiter, int(a["niter"].sum()))
