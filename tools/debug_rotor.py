"""debug: worst well-posed model of the rotor21 LVG sweep at tbg 10.926 (kernel v1 vs oracle)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_general import rotor_draws, gpu_solve_dens, ROTOR
from oracle import parity
from radex_emcee_b200 import _lib
ctx = _lib.Context(_lib.MolData(ROTOR), 0)
method, tbg = 2, 10.926
T, d, N = rotor_draws(np.random.default_rng(40 + method + int(tbg)), 300, tbg)
got = gpu_solve_dens(ctx, T, d, N, tbg, method)
ref, cls, runs = parity.classify(ROTOR, T, d, N, tbg, method, more=True)
ex, et, eu, es = parity.rel_errors(got, ref, ref["iupp"])
w = np.maximum(np.maximum(ex, et), np.maximum(eu, es))
wp = cls["well_posed"]
i = int(np.argmax(np.where(wp, w, 0)))
print("model", i, T[i], d[i], N[i], "errs pops/tex/tau/flux", ex[i], et[i], eu[i], es[i], "niter", got["niter"][i], ref["niter"][i])
print("spread of the reference under the 13 perturbations:", max(parity.worst(r, ref, ref["iupp"])[i] for r in runs))
np.set_printoptions(linewidth=200, precision=6)
print("xpop ref", ref["xpop"][i][:12]); print("xpop gpu", got["xpop"][i][:12])
print("tex ref", ref["tex"][i][:10]); print("tex gpu", got["tex"][i][:10])
print("tau ref", ref["tau"][i][:10]); print("tau gpu", got["tau"][i][:10])
for mi in (50, 100, 150, 199):
    a = gpu_solve_dens(ctx, T[i:i+1], d[i:i+1], N[i:i+1], tbg, method, maxiter=mi, abs_tol=0.0)
    from oracle.oracle import Oracle
    o = Oracle(ROTOR)
    b = o.solve_batch_dens(T[i:i+1], d[i:i+1], N[i:i+1], tbg=tbg, method=method, maxiter=mi, abs_tol=0.0)
    print("after", mi, "calls: max rel diff pops", np.max(np.abs(a["xpop"][0] - b["xpop"][0]) / b["xpop"][0]), "tex", np.max(np.abs(a["tex"][0]-b["tex"][0])/np.abs(b["tex"][0])))
