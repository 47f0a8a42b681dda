"""One-component SLED model: vectorised mirror of emcee/emcee_radex.py:95-181.

Same function names and argument meaning as the reference script; every function also accepts a
batch of parameter vectors ``p`` of shape (n, 4) and then returns n values, which is the form
``emcee.EnsembleSampler(..., vectorize=True)`` calls.  ``lnprob`` is one fused kernel launch
(prior -> solve -> line fluxes -> chi^2); ``model_lvg``/``lnlike``/``lnprior`` are provided
separately with the reference's semantics for callers that use them (curve_fit, plots).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .cosmo import r_angle
from .radex import Radex

opr = 3
fortho = opr / (1 + opr)

R = None


def init_radex(tbg=2.7315, device=0, datapath=None):
    """Create the global solver handle (emcee/emcee_radex.py:104-117)."""
    global R
    if R is None:
        R = Radex(species="co", datapath=datapath,
                  density={"oH2": fortho * 10. ** 10.0, "pH2": (1 - fortho) * 10. ** 10.0},
                  column=10.0 ** 6.0, temperature=20.0, tbackground=tbg, deltav=1.0,
                  escapeProbGeom="lvg", device=device)
    return R


def _as2d(p, ndim):
    p = np.asarray(p, dtype=np.float64)
    single = p.ndim == 1
    p2 = np.ascontiguousarray(p.reshape(-1, ndim))
    return p2, single


def model_lvg(Jup, params, R=None):
    """Model fluxes [Jy km/s] at the observed lines (emcee/emcee_radex.py:120-130)."""
    p, single = _as2d(params, 4)
    R.set_params(density={"oH2": fortho * 10. ** p[:, 0], "pH2": (1 - fortho) * 10. ** p[:, 0]},
                 column=10. ** p[:, 2], temperature=10. ** p[:, 1])
    R.run_radex(validate_colliders=False, reuse_last=True, reload_molfile=False)
    result = np.atleast_2d(R.source_line_surfbrightness)
    idx = np.asarray(np.int_(Jup)) - 1
    intensity = result[:, idx] * (10. ** p[:, 3])[:, None] * 1.0e23      # x sr x 1 km/s -> Jy km/s
    return intensity[0] if single else intensity


def lnlike(p, Jup, flux, eflux, R=None, sigma_floor=1e-12):
    """emcee/emcee_radex.py:132-167."""
    p2, single = _as2d(p, 4)
    out = np.full(p2.shape[0], -np.inf)
    flux = np.asarray(flux, dtype=np.float64)
    eflux = np.asarray(eflux, dtype=np.float64)
    # ValueError (T or N out of pyradex's range) -> -inf, per walker
    T, N = 10. ** p2[:, 1], 10. ** p2[:, 2]
    ok = (T > 0) & (T <= 1e4) & (N >= 1e5) & (N <= 1e25)
    if ok.any():
        model = np.atleast_2d(model_lvg(Jup, p2[ok], R))
        e = np.maximum(np.abs(eflux), sigma_floor)
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            r = (flux - model) / e
            max_safe = np.sqrt(np.finfo(np.float64).max) / 10.0
            good = (np.all(np.isfinite(flux)) & np.all(np.isfinite(model), axis=1) & np.all(np.isfinite(e))
                    & np.all(np.isfinite(r), axis=1) & ~np.any(np.abs(r) > max_safe, axis=1))
            val = -0.5 * (np.einsum("ij,ij->i", r, r) + 2.0 * np.sum(np.log(e)))
        res = np.where(good, val, -np.inf)
        out[ok] = res
    return out[0] if single else out


def lnprior(p, bounds, R=None):
    """emcee/emcee_radex.py:169-175."""
    p2, single = _as2d(p, 4)
    bounds = np.asarray(bounds, dtype=np.float64)
    bad = np.any(p2 > bounds[:, 1], axis=1) | np.any(p2 < bounds[:, 0], axis=1)
    d = p2[:, 2] - p2[:, 0]
    bad |= (d >= 17.5) | (d <= 10.0)
    out = np.where(bad, -np.inf, 0.0)
    return out[0] if single else out


def lnprob(p, Jup, flux, eflux, bounds=None, opts=None, return_nsolves=False):
    """emcee/emcee_radex.py:177-181, one fused launch for all rows of ``p``."""
    p2, single = _as2d(p, 4)
    obs = _lib.make_obs(Jup, flux, eflux)
    b = np.ascontiguousarray(bounds, dtype=np.float64)
    if b.shape != (4, 2):
        raise ValueError("bounds must have shape (4, 2)")
    out = np.empty(p2.shape[0])
    ns = C.c_int64(0)
    o = opts if opts is not None else _lib.default_opts()
    _lib.check(_lib.load().rb_lnprob1(R._ctx.handle, p2.shape[0], _lib.ptr(p2), C.byref(obs), _lib.ptr(b), R.tbg,
                                      C.byref(o), _lib.ptr(out), C.byref(ns)))
    res = out[0] if single else out
    return (res, ns.value) if return_nsolves else res


def source_setup(z):
    """tbg, R_angle, bounds and p0 of one source (emcee/emcee_radex.py:419-451)."""
    tbg = 2.7315 * (1 + z)
    ra = r_angle(z)
    bounds = np.array([[2.0, 7.0], [np.log10(tbg), 3.0], [15.5, 19.5], [np.log10(ra) - 4, np.log10(ra) + 4]])
    p0 = np.clip([4.0, 1.4, 17.8, -9.85], bounds[:, 0], bounds[:, 1])
    return tbg, ra, bounds, p0
