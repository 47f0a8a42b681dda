"""Small runs of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck  python tools/sanitize.py [n] [maxiter]
  compute-sanitizer --tool racecheck python tools/sanitize.py 8192 6
Covers: the scheduled solve with all five cached-engine kernels (two models per warp; two rows per lane above 16 lead
levels), launches A / B / C and the capture buffer's atomic cursor; the pipelined and the fused lnprob (several sources,
per-model background); the stretch move in its second form (pack / propose2 / accept2) and the in-library loop with its
CUDA graph; the general-molecule kernels on the second table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MOLFILE, draw_params  # noqa: E402
from radex_emcee_b200 import _lib  # noqa: E402
from test_gpu_solve import gpu_solve  # noqa: E402
from test_gpu_general import ROTOR, gpu_solve_dens, rotor_draws  # noqa: E402
from test_gpu_sampler import model1, model2  # noqa: E402
from radex_emcee_b200.sampler import CudaEngine, StretchSampler  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
P = draw_params(np.random.default_rng(5), n, 10.926)
a = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, maxiter=maxiter, park_max=7)    # all five cached-engine kernels
b = gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 10.926, maxiter=maxiter, kernel=3)
for k in a:
    assert np.array_equal(a[k], b[k], equal_nan=True), k
print("solve ok", n, maxiter, int(a["niter"].sum()))

# general-molecule kernels
rctx = _lib.Context(_lib.MolData(ROTOR), 0)
T, d, N = rotor_draws(np.random.default_rng(1), 64, 2.7315)
r = gpu_solve_dens(rctx, T, d, N, 2.7315, 2, maxiter=maxiter)
print("v1 ok", int(r["niter"].sum()))

# sampler: several sources in one ensemble (fused lnprob, per-model background), in-library loop with its CUDA graph,
# and the per-half-step calls; a two-component ensemble large enough for the lnprob pipeline
opts = _lib.default_opts(maxiter=maxiter)
models, starts = [], []
from radex_emcee_b200.data import read_data  # noqa: E402
for k, nm in enumerate(list(read_data(ROOT + "/data/flux.dat"))[:3]):
    m, p0 = model1(nm)
    m.opts = opts
    models.append(m)
    starts.append(p0 + 1e-3 * np.random.default_rng(k).standard_normal((32, 4)))
for native in (True, False):
    s = StretchSampler(96, 4, CudaEngine(ctx, models), seed=3, nsources=3, native=native)
    s.run_mcmc(np.vstack(starts), 5)
    print("sampler ok native=%s" % native, s.get_chain().shape, float(np.mean(s.acceptance_fraction)))
m2, p2 = model2()
m2.opts = opts
s = StretchSampler(16384, 8, CudaEngine(ctx, m2), seed=4, native=True)
s.run_mcmc(p2 + 0.05 * np.random.default_rng(9).standard_normal((16384, 8)), 2)
print("pipeline sampler ok", s.total_solves)
