"""ctypes front-end of oracle/radex_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  See radex_oracle.c for the parity status of each routine.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libradex_oracle.so")

STOP_PYRADEX, STOP_RADEX = 0, 1
# astropy (CODATA 2018) values of h c / k_B and 2 h c in cgs: what core.py:981-984 evaluates to
FK_ASTROPY = 1.4387768775039338
THC_ASTROPY = 3.9728917142978115e-16
# RADEX's own radex.inc constants
FK_RADEX = 1.4387809925261357
THC_RADEX = 3.972907393443411e-16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    src = os.path.join(_HERE, "radex_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libradex_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ro_mol_load.restype = C.c_void_p
        L.ro_mol_load.argtypes = [C.c_char_p]
        L.ro_mol_free.argtypes = [C.c_void_p]
        for f in ("ro_mol_nlev", "ro_mol_nline", "ro_mol_npart"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ro_mol_get_levels.argtypes = [C.c_void_p, _dp, _dp]
        L.ro_mol_get_lines.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp, _dp]
        L.ro_state_new.restype = C.c_void_p
        L.ro_state_new.argtypes = [C.c_void_p]
        L.ro_state_free.argtypes = [C.c_void_p]
        for f in ("ro_state_xpop", "ro_state_tex", "ro_state_taul", "ro_state_backi", "ro_state_totalb",
                  "ro_state_crate", "ro_state_ctot"):
            getattr(L, f).restype = _dp
            getattr(L, f).argtypes = [C.c_void_p]
        L.ro_state_totdens.restype = C.c_double
        L.ro_state_totdens.argtypes = [C.c_void_p]
        L.ro_state_set_method.argtypes = [C.c_void_p, C.c_int]
        L.ro_state_set_column.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ro_set_physics.argtypes = [C.c_void_p, C.c_double, _dp, C.c_int]
        L.ro_backrad.argtypes = [C.c_void_p, C.c_double]
        L.ro_escprob.restype = C.c_double
        L.ro_escprob.argtypes = [C.c_double, C.c_int]
        L.ro_matrix.argtypes = [C.c_void_p, C.c_int]
        L.ro_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.ro_surface_brightness.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp]
        L.ro_solve.argtypes = [C.c_void_p] + [C.c_double] * 6 + [C.c_int] * 4 + [C.c_double] * 3 + [_dp, _ip]
        L.ro_lnlike.restype = C.c_double
        L.ro_lnlike.argtypes = [_dp, _dp, _dp, C.c_int]
        L.ro_lnprior1.restype = C.c_double
        L.ro_lnprior1.argtypes = [_dp, _dp]
        L.ro_lnprior2.restype = C.c_double
        L.ro_lnprior2.argtypes = [_dp, _dp, C.c_int, C.c_double]
        L.ro_lnprob1.restype = C.c_double
        L.ro_lnprob1.argtypes = [C.c_void_p, _dp, _ip, _dp, _dp, C.c_int, _dp, C.c_double, C.c_int, C.c_int,
                                 C.c_int, C.c_double, C.c_double, C.c_double]
        L.ro_lnprob2.restype = C.c_double
        L.ro_lnprob2.argtypes = [C.c_void_p, _dp, _ip, _dp, _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double,
                                 C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ro_solve_batch.argtypes = [C.c_void_p, C.c_long, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                     _dp, _dp, _dp, _dp, _ip, _ip]
        L.ro_solve_batch_dens.argtypes = [C.c_void_p, C.c_long, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                          _dp, _dp, _dp, _dp, _ip, _ip]
        L.ro_mol_get_partner_ids.argtypes = [C.c_void_p, _ip]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """One RADEX COMMON-block state + the pyradex loop around it (history carried like the reference)."""

    def __init__(self, molfile):
        self.L = lib()
        self.mol = self.L.ro_mol_load(os.fsencode(molfile))
        if not self.mol:
            raise ValueError("cannot parse LAMDA file %s" % molfile)
        self.nlev = self.L.ro_mol_nlev(self.mol)
        self.nline = self.L.ro_mol_nline(self.mol)
        self.st = self.L.ro_state_new(self.mol)
        self.eterm = np.zeros(self.nlev)
        self.gstat = np.zeros(self.nlev)
        self.L.ro_mol_get_levels(self.mol, _d(self.eterm), _d(self.gstat))
        self.iupp = np.zeros(self.nline, np.int32)
        self.ilow = np.zeros(self.nline, np.int32)
        self.aeinst, self.spfreq, self.eup, self.xnu = (np.zeros(self.nline) for _ in range(4))
        self.L.ro_mol_get_lines(self.mol, _i(self.iupp), _i(self.ilow), _d(self.aeinst), _d(self.spfreq),
                                _d(self.eup), _d(self.xnu))

    def __del__(self):
        try:
            self.L.ro_state_free(self.st)
            self.L.ro_mol_free(self.mol)
        except Exception:
            pass

    def _view(self, fn, n):
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.st), shape=(n,))

    @property
    def xpop(self):
        return self._view("ro_state_xpop", self.nlev)

    @property
    def tex(self):
        return self._view("ro_state_tex", self.nline)

    @property
    def taul(self):
        return self._view("ro_state_taul", self.nline)

    @property
    def backi(self):
        return self._view("ro_state_backi", self.nline)

    @property
    def totalb(self):
        return self._view("ro_state_totalb", self.nline)

    @property
    def crate(self):
        return self._view("ro_state_crate", self.nlev * self.nlev).reshape(self.nlev, self.nlev)

    @property
    def ctot(self):
        return self._view("ro_state_ctot", self.nlev)

    @property
    def totdens(self):
        return self.L.ro_state_totdens(self.st)

    def set_physics(self, tkin, n_ph2, n_oh2, n_h2=0.0):
        d = np.array([n_h2, n_ph2, n_oh2, 0, 0, 0, 0], dtype=np.float64)
        self.L.ro_set_physics(self.st, float(tkin), _d(d), 7)

    def set_physics_dens(self, tkin, dens7):
        d = np.ascontiguousarray(dens7, dtype=np.float64)
        assert d.shape == (7,)
        self.L.ro_set_physics(self.st, float(tkin), _d(d), 7)

    def set_column(self, cdmol, deltav_kms=1.0):
        self.L.ro_state_set_column(self.st, float(cdmol), float(deltav_kms) * 1e5)

    def set_method(self, method):
        self.L.ro_state_set_method(self.st, int(method))

    def backrad(self, tbg):
        self.L.ro_backrad(self.st, float(tbg))

    def escprob(self, tau, method=2):
        return self.L.ro_escprob(float(tau), int(method))

    def matrix(self, niter):
        return self.L.ro_matrix(self.st, int(niter))

    def run(self, reuse_last=False, stop_rule=STOP_PYRADEX, miniter=10, maxiter=200, abs_tol=1e-16):
        return self.L.ro_run(self.st, int(reuse_last), stop_rule, miniter, maxiter, abs_tol)

    def surface_brightness(self, fk=FK_ASTROPY, thc=THC_ASTROPY):
        out = np.zeros(self.nline)
        self.L.ro_surface_brightness(self.st, fk, thc, _d(out))
        return out

    def solve_batch(self, tkin, n_ph2, n_oh2, cdmol, deltav_kms=1.0, tbg=2.7315, method=2,
                    stop_rule=STOP_PYRADEX, miniter=10, maxiter=200, abs_tol=1e-16,
                    fk=FK_ASTROPY, thc=THC_ASTROPY):
        tkin, n_ph2, n_oh2, cdmol = (np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
                                     for x in (tkin, n_ph2, n_oh2, cdmol))
        n = tkin.size
        out = dict(xpop=np.zeros((n, self.nlev)), tex=np.zeros((n, self.nline)), tau=np.zeros((n, self.nline)),
                   surf=np.zeros((n, self.nline)), niter=np.zeros(n, np.int32), status=np.zeros(n, np.int32))
        self.L.ro_solve_batch(self.st, n, _d(tkin), _d(n_ph2), _d(n_oh2), _d(cdmol), deltav_kms, tbg, method,
                              stop_rule, miniter, maxiter, abs_tol, fk, thc, _d(out["xpop"]), _d(out["tex"]),
                              _d(out["tau"]), _d(out["surf"]), _i(out["niter"]), _i(out["status"]))
        return out

    def solve_batch_dens(self, tkin, dens7, cdmol, deltav_kms=1.0, tbg=2.7315, method=2,
                         stop_rule=STOP_PYRADEX, miniter=10, maxiter=200, abs_tol=1e-16,
                         fk=FK_ASTROPY, thc=THC_ASTROPY):
        """dens7[n, 7]: one density per LAMDA partner id 1..7 (H2, p-H2, o-H2, e, H, He, H+), as pyradex writes
        cphys.density (core.py:525-561)."""
        tkin, cdmol = (np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64) for x in (tkin, cdmol))
        n = tkin.size
        dens7 = np.ascontiguousarray(np.broadcast_to(np.asarray(dens7, dtype=np.float64), (n, 7)))
        out = dict(xpop=np.zeros((n, self.nlev)), tex=np.zeros((n, self.nline)), tau=np.zeros((n, self.nline)),
                   surf=np.zeros((n, self.nline)), niter=np.zeros(n, np.int32), status=np.zeros(n, np.int32))
        self.L.ro_solve_batch_dens(self.st, n, _d(tkin), _d(dens7), _d(cdmol), deltav_kms, tbg, method, stop_rule,
                                   miniter, maxiter, abs_tol, fk, thc, _d(out["xpop"]), _d(out["tex"]), _d(out["tau"]),
                                   _d(out["surf"]), _i(out["niter"]), _i(out["status"]))
        return out

    @property
    def partner_ids(self):
        ids = np.zeros(self.L.ro_mol_npart(self.mol), np.int32)
        self.L.ro_mol_get_partner_ids(self.mol, _i(ids))
        return ids

    def lnprob1(self, p, jup, flux, eflux, bounds, tbg, stop_rule=STOP_PYRADEX, miniter=10, maxiter=200,
                abs_tol=1e-16, fk=FK_ASTROPY, thc=THC_ASTROPY):
        p = np.ascontiguousarray(p, dtype=np.float64)
        jup = np.ascontiguousarray(jup, dtype=np.int32)
        flux = np.ascontiguousarray(flux, dtype=np.float64)
        eflux = np.ascontiguousarray(eflux, dtype=np.float64)
        bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        return self.L.ro_lnprob1(self.st, _d(p), _i(jup), _d(flux), _d(eflux), jup.size, _d(bounds), tbg,
                                 stop_rule, miniter, maxiter, abs_tol, fk, thc)

    def lnprob2(self, p, jup, flux, eflux, bounds, t_d, tbg, stop_rule=STOP_PYRADEX, miniter=10, maxiter=200,
                abs_tol=1e-16, fk=FK_ASTROPY, thc=THC_ASTROPY):
        p = np.ascontiguousarray(p, dtype=np.float64)
        jup = np.ascontiguousarray(jup, dtype=np.int32)
        flux = np.ascontiguousarray(flux, dtype=np.float64)
        eflux = np.ascontiguousarray(eflux, dtype=np.float64)
        bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        has_td = t_d is not None
        return self.L.ro_lnprob2(self.st, _d(p), _i(jup), _d(flux), _d(eflux), jup.size, _d(bounds), int(has_td),
                                 float(t_d) if has_td else 0.0, tbg, stop_rule, miniter, maxiter, abs_tol, fk, thc)
