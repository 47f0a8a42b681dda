#!/usr/bin/env python
"""what each stop rule costs and what it leaves unconverged.  For the config-2 sweep (2^log2n models):
the pyradex rule at abs_convergence_threshold 1e-16 (the reference's default: the fixed point), at looser thresholds
(run_radex exposes the argument, core.py:857), and RADEX's own conv flag; per rule: iterations per solve, solves/s
(CUDA events, one warm-up), and the error of populations and fluxes of every model against the fixed point.
  python tools/tol_sweep.py --log2n 18 > gpurun_out/tol_sweep.jsonl"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import draw, TBG, MOLFILE
from radex_emcee_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=18)
args = ap.parse_args()
n = 1 << args.log2n
dev = torch.device("cuda", 0)
mol = _lib.MolData(MOLFILE)
ctx = _lib.Context(mol, 0)
L = _lib.load()
tk, nh2, cd = draw(n, 0)
dens = np.zeros((n, mol.npart))
for p, pid in enumerate(mol.partner_id):
    dens[:, p] = {2: 0.25, 3: 0.75}.get(int(pid), 0.0) * nh2
d_tk, d_dens, d_cd = (torch.from_numpy(a).to(dev) for a in (tk, dens, cd))
stream = torch.cuda.current_stream(dev)
ctx.set_stream(stream.cuda_stream)


def run(**kw):
    o = _lib.default_opts(**kw)
    x = torch.empty((n, mol.nlev), dtype=torch.float64, device=dev)
    s = torch.empty((n, mol.nline), dtype=torch.float64, device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    ms = None
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.rb_solve_batch_dev(ctx.handle, n, d_tk.data_ptr(), d_dens.data_ptr(), d_cd.data_ptr(), 1.0, TBG, 2,
                                        C.byref(o), x.data_ptr(), None, None, s.data_ptr(), it.data_ptr(), st.data_ptr()))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    return x, s, it, st, ms, ctx.counters()[0]


fx, fs, fit, fst, fms, fiters = run()
qs = torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.float64, device=dev)
sig = fx > 1e-9
bright = (fs.abs() > 1e-6 * torch.nan_to_num(fs.abs(), nan=0.0).amax(dim=1, keepdim=True)) & (fs.abs() > 1e-25)
conv = (fst & 4) == 0        # the fixed point is only defined where the reference's rule stopped
rules = [("pyradex abs_tol=1e-16 (reference default)", {}), ("pyradex abs_tol=1e-14", {"abs_tol": 1e-14}),
         ("pyradex abs_tol=1e-12", {"abs_tol": 1e-12}), ("pyradex abs_tol=1e-10", {"abs_tol": 1e-10}),
         ("pyradex abs_tol=1e-8", {"abs_tol": 1e-8}), ("pyradex abs_tol=1e-6", {"abs_tol": 1e-6}),
         ("RADEX conv flag", {"stop_rule": 1})]
for label, kw in rules:
    x, s, it, st, ms, iters = run(**kw) if kw else (fx, fs, fit, fst, fms, fiters)
    ex = torch.where(sig, (x - fx).abs() / fx, torch.zeros_like(fx)).amax(dim=1)
    es = torch.where(bright, (s - fs).abs() / fs.abs(), torch.zeros_like(fs)).amax(dim=1)
    ok = conv & torch.isfinite(ex) & torch.isfinite(es)

    def quant(v):
        v = torch.sort(v[ok]).values
        return [float(a) for a in v[(qs * (v.numel() - 1)).long()]] + [float(v[-1])]
    print(json.dumps({"rule": label, "models": n, "compared": int(ok.sum()), "iters_per_solve": iters / n,
                      "solves_per_s": n / (ms * 1e-3), "ms": ms, "frac_at_maxiter": float(((st & 4) != 0).double().mean()),
                      "quantiles": [0.5, 0.9, 0.99, 0.999, 1.0], "pops_rel_err": quant(ex), "flux_rel_err": quant(es),
                      "frac_flux_within_1e-5": float((es[ok] < 1e-5).double().mean()),
                      "frac_pops_within_1e-5": float((ex[ok] < 1e-5).double().mean())}))
