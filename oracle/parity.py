"""Parity bookkeeping shared by tests/, __graft_entry__.smoke() and bench.py's `parity` record -- TEST INFRASTRUCTURE.

What is compared (BASELINE.json north_star: populations and line fluxes within 1e-5 relative): per model the largest
relative error of the populations the reference resolves (> 1e-9), of Tex / tau of those levels' lines, and of the
fluxes of lines brighter than 1e-6 of the model's brightest AND standing out from the background they absorb by more
than 0.1 % (`contrast`): the brightness is toti - backi = (B(Tex) - B(Tbg))(1 - e^-tau), and where Tex sits within ~1e-3
of Tbg that difference cancels -- the reference's own value moves by 1e-4 under a 1e-8 change of an input while its
populations, Tex and tau stay put to 1e-10 (measured: rotor21, T = 11.6 K against Tbg = 10.9 K).

Where it is compared.  The under-relaxed RADEX iteration is not a contraction everywhere: some models have several
attractors or end in a limit cycle at maxiter, and there the REFERENCE'S OWN answer changes by O(1) when an input
moves in its 13th digit.  No implementation with different rounding can match it there, so models are classified
with the oracle alone (never with the GPU result):
  nonfinite   the oracle's own brightness is NaN/inf (LVG escape probability for tau <= -14, ...)
  maser       a line with tau < -3 (amplification e^-tau: hypersensitive by construction)
  sensitive   the oracle re-run with one input perturbed by 3e-14 .. 1e-11 (PERTURBATIONS) moves by more than 1e-6
  well_posed  none of the above: the 1e-5 bar applies
For the excluded classes the checkable statement is that the GPU lands on one of the reference's own answers
(`attractor_error`: distance to the nearest of the unperturbed and perturbed oracle runs).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .oracle import Oracle

PERTURBATIONS = ((0, 3e-14), (1, 1e-13), (2, 1e-13), (0, -1e-12), (2, -1e-11))
# a wider fan for the attractor question only (which answers does the reference itself give near this input?)
MORE_PERTURBATIONS = ((1, -3e-13), (0, 1e-11), (2, 3e-12), (1, 1e-10), (0, -1e-10), (2, 1e-9), (1, -1e-9), (0, 1e-8))


def solve_threads(molfile, T, nh2, N, tbg, method=2, nthreads=None, **kw):
    """Oracle.solve_batch over the host cores (one RADEX COMMON-block state per thread; ctypes drops the GIL).
    Every solve starts clean, so the result does not depend on the split.  nh2: total n(H2), split 1:3 into p-/o-H2
    like the drivers (opr = 3) -- or an [n, 7] array with one density per LAMDA partner id (H2, p-H2, o-H2, e, H, He, H+)."""
    T, N = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (T, N))
    nh2 = np.ascontiguousarray(nh2, dtype=np.float64)
    if nh2.ndim == 2:
        return _solve_threads_dens(molfile, T, nh2, N, tbg, method, nthreads, **kw)
    nh2 = np.atleast_1d(nh2)
    n = T.size
    nthreads = max(1, min(nthreads or (os.cpu_count() or 1), 64, n))
    chunks = np.array_split(np.arange(n), nthreads)
    oracles = [Oracle(molfile) for _ in range(nthreads)]

    def work(t):
        i = chunks[t]
        return oracles[t].solve_batch(T[i], 0.25 * nh2[i], 0.75 * nh2[i], N[i], tbg=tbg, method=method, **kw)

    with ThreadPoolExecutor(nthreads) as ex:
        parts = list(ex.map(work, range(nthreads)))
    out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    out["iupp"] = oracles[0].iupp.copy()
    _add_contrast(out, oracles[0], tbg)
    return out


def _add_contrast(out, o, tbg):
    """|toti - backi| / (backi (1 - e^-tau)): how far a line stands out from the background it absorbs."""
    o.backrad(float(tbg))
    with np.errstate(all="ignore"):
        absorbed = o.backi[None, :] * np.abs(1.0 - np.exp(-out["tau"]))
        out["contrast"] = np.abs(out["surf"]) / np.maximum(absorbed, 1e-300)


def _solve_threads_dens(molfile, T, dens7, N, tbg, method, nthreads, **kw):
    n = T.size
    nthreads = max(1, min(nthreads or (os.cpu_count() or 1), 64, n))
    chunks = np.array_split(np.arange(n), nthreads)
    oracles = [Oracle(molfile) for _ in range(nthreads)]

    def work(t):
        i = chunks[t]
        return oracles[t].solve_batch_dens(T[i], dens7[i], N[i], tbg=tbg, method=method, **kw)

    with ThreadPoolExecutor(nthreads) as ex:
        parts = list(ex.map(work, range(nthreads)))
    out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    out["iupp"] = oracles[0].iupp.copy()
    _add_contrast(out, oracles[0], tbg)
    return out


def rel_errors(got, ref, iupp):
    """Per-model max relative errors (populations, Tex, tau, flux) on the entries the reference resolves."""
    with np.errstate(all="ignore"):
        xr = ref["xpop"]
        sig = xr > 1e-9
        ex = np.where(sig, np.abs(got["xpop"] - xr) / xr, 0).max(axis=1)
        sl = sig[:, iupp - 1]
        et = np.where(sl, np.abs(got["tex"] - ref["tex"]) / np.abs(ref["tex"]), 0)
        eu = np.where(sl, np.abs(got["tau"] - ref["tau"]) / np.maximum(np.abs(ref["tau"]), 1e-12), 0)
        sr = ref["surf"]
        bright = np.abs(sr) > 1e-6 * np.nanmax(np.abs(sr), axis=1, keepdims=True)
        bright &= np.abs(sr) > 1e-25          # erg s-1 cm-2 Hz-1 sr-1; real lines are 1e-16 .. 1e-9
        if "contrast" in ref:
            bright &= ref["contrast"] > 1e-3  # not a difference of two nearly equal brightnesses
        es = np.where(bright & sl, np.abs(got["surf"] - sr) / np.abs(sr), 0)
    f = lambda e: np.nan_to_num(e, nan=np.inf).max(axis=1)
    return np.nan_to_num(ex, nan=np.inf), f(et), f(eu), f(es)


def worst(got, ref, iupp):
    """max over (populations, Tex, tau, flux) per model."""
    return np.max(np.vstack(rel_errors(got, ref, iupp)), axis=0)


def classify(molfile, T, nh2, N, tbg, method=2, ref=None, nthreads=None, more=False, **kw):
    """Oracle-only classification.  Returns (ref, classes, runs): classes maps name -> bool mask (exclusive, in the
    order nonfinite, maser, sensitive, well_posed); runs = the perturbed oracle results (for attractor_error)."""
    T, N = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (T, N))
    nh2 = np.ascontiguousarray(nh2, dtype=np.float64)      # [n] total n(H2), or [n, 7] per partner id
    if ref is None:
        ref = solve_threads(molfile, T, nh2, N, tbg, method, nthreads, **kw)
    iupp = ref["iupp"] if "iupp" in ref else Oracle(molfile).iupp
    nonfinite = ~np.isfinite(ref["surf"]).all(axis=1)
    maser = ~nonfinite & ~(np.nan_to_num(ref["tau"], nan=-np.inf).min(axis=1) > -3.0)
    moved = np.zeros(T.size, dtype=bool)
    runs = []
    for which, eps in PERTURBATIONS + (MORE_PERTURBATIONS if more else ()):
        t, d, c = T.copy(), nh2.copy(), N.copy()
        (t, d, c)[which][:] *= 1 + eps
        pert = solve_threads(molfile, t, d, c, tbg, method, nthreads, **kw)
        runs.append(pert)
        if (which, eps) in PERTURBATIONS:
            ex, et, eu, es = rel_errors(pert, ref, iupp)
            moved |= ~((ex < 1e-6) & (et < 1e-6) & (es < 1e-6))
    sensitive = ~nonfinite & ~maser & moved
    classes = {"nonfinite": nonfinite, "maser": maser, "sensitive": sensitive,
               "well_posed": ~nonfinite & ~maser & ~sensitive}
    return ref, classes, runs


def attractor_error(got, ref, runs, iupp):
    """Per model: the smallest `worst` error of `got` against the reference's own answers (unperturbed + perturbed)."""
    best = worst(got, ref, iupp)
    for r in runs:
        best = np.minimum(best, worst(got, r, iupp))
    return best


def summary(got, molfile, T, nh2, N, tbg, method=2, nthreads=None, more=True, **kw):
    """The record bench.py prints and tests assert on: error statistics on ALL models and per class."""
    ref, cls, runs = classify(molfile, T, nh2, N, tbg, method, nthreads=nthreads, more=more, **kw)
    iupp = ref["iupp"]
    ex, et, eu, es = rel_errors(got, ref, iupp)
    w = np.maximum(np.maximum(ex, et), np.maximum(eu, es))
    att = attractor_error(got, ref, runs, iupp)
    wp = cls["well_posed"]
    fin = np.isfinite(w)
    q = lambda a, m: float(np.max(a[m])) if m.any() else None
    med = lambda a, m: float(np.median(a[m])) if m.any() else None
    rec = {
        "models": int(T.size),
        "tolerance": 1e-5,
        "classes": {k: int(v.sum()) for k, v in cls.items()},
        "well_posed_fraction": float(wp.mean()),
        "well_posed": {"max_rel_err_pops": q(ex, wp), "max_rel_err_flux": q(es, wp), "max_rel_err_tex": q(et, wp),
                       "max_rel_err_tau": q(eu, wp), "median_rel_err_pops": med(ex, wp),
                       "median_rel_err_flux": med(es, wp), "within_tolerance": int((w[wp] < 1e-5).sum())},
        "all_models": {"median_rel_err_pops": med(ex, fin), "median_rel_err_flux": med(es, fin),
                       "max_rel_err_pops": q(ex, fin), "max_rel_err_flux": q(es, fin),
                       "within_tolerance": int((w < 1e-5).sum()), "nonfinite_error": int((~fin).sum())},
        "excluded": {k: {"models": int(cls[k].sum()),
                         "within_tolerance_of_reference": int((w[cls[k]] < 1e-5).sum()),
                         "within_tolerance_of_a_reference_attractor": int((att[cls[k]] < 1e-5).sum())}
                     for k in ("nonfinite", "maser", "sensitive")},
        "niter": {"median_abs_diff": float(np.median(np.abs(got["niter"] - ref["niter"]))),
                  "equal": int((got["niter"] == ref["niter"]).sum()),
                  "reference_at_maxiter": int((ref["status"] & 4).astype(bool).sum()),
                  "gpu_at_maxiter": int((np.asarray(got["status"]) & 4).astype(bool).sum())},
        "perturbations": len(runs),
    }
    return rec, ref, cls, w, att
