"""Two-component (cold + warm) SLED model: vectorised mirror of emcee/emcee_radex_2comp.py:99-244."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import emcee_radex as _one
from .cosmo import r_angle

opr = 3.0
fortho = opr / (1.0 + opr)

R = None


def init_radex(tbg=2.7315, device=0, datapath=None):
    """emcee/emcee_radex_2comp.py:106-119."""
    global R
    if R is None:
        saved, _one.R = _one.R, None
        R = _one.init_radex(tbg, device=device, datapath=datapath)
        _one.R = saved
    return R


def _as2d(p, ndim):
    p = np.asarray(p, dtype=np.float64)
    single = p.ndim == 1
    return np.ascontiguousarray(p.reshape(-1, ndim)), single


def model_single_lvg(Jup, p, R=None):
    """emcee/emcee_radex_2comp.py:150-162."""
    return _one.model_lvg(Jup, p, R)


def model_lvg(Jup, p, R=None):
    """Sum of two solves with two sizes (emcee/emcee_radex_2comp.py:122-147)."""
    p2, single = _as2d(p, 8)
    a = np.atleast_2d(_one.model_lvg(Jup, p2[:, :4], R))
    b = np.atleast_2d(_one.model_lvg(Jup, p2[:, 4:], R))
    out = a + b
    return out[0] if single else out


def residual(p, R=None, Jup=None, flux=None, eflux=None):
    return (flux - model_lvg(Jup, p, R)) / eflux


def lnlike(p, Jup, flux, eflux, R=None, sigma_floor=1e-12):
    """emcee/emcee_radex_2comp.py:169-196."""
    p2, single = _as2d(p, 8)
    out = np.full(p2.shape[0], -np.inf)
    flux = np.asarray(flux, dtype=np.float64)
    eflux = np.asarray(eflux, dtype=np.float64)
    T = 10. ** p2[:, [1, 5]]
    N = 10. ** p2[:, [2, 6]]
    ok = np.all((T > 0) & (T <= 1e4) & (N >= 1e5) & (N <= 1e25), axis=1)
    if ok.any():
        model = np.atleast_2d(model_lvg(Jup, p2[ok], R))
        e = np.maximum(np.abs(eflux), sigma_floor)
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            r = (flux - model) / e
            max_safe = np.sqrt(np.finfo(np.float64).max) / 10.0
            good = (np.all(np.isfinite(flux)) & np.all(np.isfinite(model), axis=1) & np.all(np.isfinite(e))
                    & np.all(np.isfinite(r), axis=1) & ~np.any(np.abs(r) > max_safe, axis=1))
            val = -0.5 * (np.einsum("ij,ij->i", r, r) + 2.0 * np.sum(np.log(e)))
        out[ok] = np.where(good, val, -np.inf)
    return out[0] if single else out


def lnprior(p, bounds, T_d=None, R=None):
    """emcee/emcee_radex_2comp.py:199-234 (note: flat terms add minus the *width* of each bound)."""
    p2, single = _as2d(p, 8)
    bounds = np.asarray(bounds, dtype=np.float64)
    bad = np.any(p2 > bounds[:, 1], axis=1) | np.any(p2 < bounds[:, 0], axis=1)
    bad |= p2[:, 5] <= p2[:, 1]
    d1, d2 = p2[:, 2] - p2[:, 0], p2[:, 6] - p2[:, 4]
    bad |= (d1 >= 18.0) | (d1 <= 9.0) | (d2 >= 18.0) | (d2 <= 9.0)
    bad |= p2[:, 3] < p2[:, 7]
    logp = np.zeros(p2.shape[0])
    for idx in range(8):
        if idx == 1 and T_d is not None:
            if T_d <= 0:
                bad |= True
                continue
            T_kin = 10.0 ** p2[:, idx]
            sigma = 1.0 * T_d
            logp = logp + (-0.5 * ((T_kin - T_d) / sigma) ** 2.0 - np.log(sigma * np.sqrt(2.0 * np.pi)))
        else:
            logp = logp + -(bounds[idx, 1] - bounds[idx, 0])
    out = np.where(bad, -np.inf, logp)
    return out[0] if single else out


def lnprob(p, Jup, flux, eflux, bounds=None, T_d=None, opts=None, return_nsolves=False):
    """emcee/emcee_radex_2comp.py:237-244, one fused launch for all rows of ``p``."""
    p2, single = _as2d(p, 8)
    obs = _lib.make_obs(Jup, flux, eflux)
    b = np.ascontiguousarray(bounds, dtype=np.float64)
    if b.shape != (8, 2):
        raise ValueError("bounds must have shape (8, 2)")
    out = np.empty(p2.shape[0])
    ns = C.c_int64(0)
    o = opts if opts is not None else _lib.default_opts()
    _lib.check(_lib.load().rb_lnprob2(R._ctx.handle, p2.shape[0], _lib.ptr(p2), C.byref(obs), _lib.ptr(b),
                                      int(T_d is not None), float(T_d) if T_d is not None else 0.0, R.tbg,
                                      C.byref(o), _lib.ptr(out), C.byref(ns)))
    res = out[0] if single else out
    return (res, ns.value) if return_nsolves else res


def source_setup(z):
    """tbg, R_angle, bounds and p0 of one source (emcee/emcee_radex_2comp.py:490-522)."""
    tbg = 2.7315 * (1 + z)
    ra = r_angle(z)
    lo, hi = np.log10(ra) - 9, np.log10(ra) + 9
    bounds = np.array([[1.5, 7.0], [np.log10(tbg), 3.0], [14.5, 19.5], [lo, hi],
                       [1.5, 7.0], [np.log10(tbg), 3.0], [14.5, 19.5], [lo, hi]])
    p0 = np.array([1.9, 1.2, 16.4, -12.1, 3.9, 2.5, 17.5, -12.1])
    return tbg, ra, bounds, p0
