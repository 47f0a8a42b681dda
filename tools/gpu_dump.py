"""Dev tool (runs on the GPU box): solve config-2 style draws on the GPU and with the oracle and dump both
to gpurun_out/ for offline analysis."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MOLFILE, draw_params
from radex_emcee_b200 import _lib
from oracle.oracle import Oracle
from test_gpu_solve import gpu_solve

kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
o = Oracle(MOLFILE)
out = {}
for method, tbg in ((2, 10.926), (1, 2.7315), (3, 10.926)):
    P = draw_params(np.random.default_rng(1000 + method + int(tbg)), n, tbg)
    T, nh2, N = P[:, 0], P[:, 1], P[:, 2]
    t0 = time.time()
    ref = o.solve_batch(T, 0.25 * nh2, 0.75 * nh2, N, tbg=tbg, method=method)
    t1 = time.time()
    got = gpu_solve(ctx, T, nh2, N, tbg, method, kernel=kernel)
    t2 = time.time()
    print("method", method, "oracle %.2fs gpu %.3fs" % (t1 - t0, t2 - t1), "niter equal frac", (got["niter"] == ref["niter"]).mean())
    out["P_%d" % method] = P
    for k in got:
        out["got_%s_%d" % (k, method)] = got[k]
        out["ref_%s_%d" % (k, method)] = ref[k]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_k%d.npz" % kernel), **out)
