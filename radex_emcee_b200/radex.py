"""`Radex`: host-side mirror of ``pyradex.Radex`` (reference: emcee/pyradex/core.py:195-1091,
emcee/pyradex/base_class.py) on top of the B200 kernels.

Same constructor/`set_params`/`run_radex` keywords, same physical units, same ``ValueError``s --
but every scalar input may also be a 1-D array of length n (a batch of models solved in one
launch), and results are plain numpy arrays (the reference returns astropy Quantities; units are
stated in each docstring).  Each solve is history-free (clean ``niter=0`` start): the reference's
``reuse_last=True`` carries the previous call's populations only as a starting guess for the same
fixed point, so the flag is accepted and ignored.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from . import _lib
from ._lib import GEOM, STOP_PYRADEX

# radex.inc constants (the binary's constant pool) -- used for the background only
FK_RADEX = 1.4387809925261357
THC_RADEX = 3.972907393443411e-16
PC_CM = 3.0856775814913673e18          # 1 pc, the hard-coded length scale (core.py:823-826)
KB_CGS = 1.380649e-16
C_CGS = 2.99792458e10

_COLLIDER_IDS = {"H2": 1, "PH2": 2, "OH2": 3, "E": 4, "H": 5, "HE": 6, "H+": 7}   # core.py:492-498
_CANON = {"H2": "H2", "PH2": "pH2", "OH2": "oH2", "E": "e", "H": "H", "HE": "He", "H+": "H+"}  # core.py:465-471


def default_datapath() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _unitless(x):
    """Strip an astropy-like Quantity (core.py uses utils.unitless)."""
    return x.value if hasattr(x, "value") else x


def is_synthetic_table(molpath: str) -> bool:
    """True for the CO-like table synth_lamda.py writes (its !MOLECULE line says so): good for tests, not for science."""
    try:
        with open(molpath, "r", errors="replace") as f:
            head = [next(f, "") for _ in range(3)]
    except OSError:
        return False
    return any("synthetic" in ln.lower() for ln in head)


_mol_cache: dict = {}
_ctx_cache: dict = {}


def get_moldata(molpath: str) -> _lib.MolData:
    """Process-wide cache: the LAMDA file is parsed once (the reference re-reads it twice per solve)."""
    key = os.path.abspath(molpath)
    if key not in _mol_cache:
        _mol_cache[key] = _lib.MolData(key)
    return _mol_cache[key]


def get_context(molpath: str, device: int = 0) -> _lib.Context:
    """Process-wide cache: tables are uploaded once per (file, device)."""
    key = (os.path.abspath(molpath), int(device))
    if key not in _ctx_cache:
        _ctx_cache[key] = _lib.Context(get_moldata(key[0]), device)
    return _ctx_cache[key]


class Radex:
    """Batched RADEX escape-probability solver with pyradex's ``Radex`` surface."""

    def __call__(self, return_table=True, **kwargs):
        self.set_params(**kwargs)
        niter = self.run_radex(reload_molfile=False, validate_colliders=False)
        return self.get_table() if return_table else niter

    def __init__(self, collider_densities=None, density=None, total_density=None, temperature=None,
                 species="co", column=None, column_per_bin=None, tbackground=2.7315, deltav=1.0,
                 abundance=None, datapath=None, escapeProbGeom="lvg", outfile="radex.out",
                 logfile="radex.log", debug=False, mu=2.8, source_area=None, device=0):
        self.mu = mu
        self.device = device
        if os.getenv("RADEX_DATAPATH") and datapath is None:
            datapath = os.getenv("RADEX_DATAPATH")
        if datapath is None:
            datapath = default_datapath()
        self.datapath = os.path.expanduser(datapath)
        self._species = None
        self.species = species

        if sum(x is not None for x in (collider_densities, density, total_density)) > 1:
            raise ValueError("Can only specify one of density, total_density, and collider_densities")
        if sum(x is not None for x in (column, column_per_bin)) > 1:
            raise ValueError("Can only specify one of column, column_per_bin.")
        n_spec = sum(x is not None for x in (column, column_per_bin, collider_densities, density,
                                             total_density, abundance))
        if n_spec > 2:
            raise ValueError("Can only specify two of column, density, and abundance.")
        if n_spec < 2:
            raise ValueError("Must specify two of column, density, and abundance.")

        self.outfile, self.logfile, self.debug = outfile, logfile, debug
        self.miniter, self.maxiter = 10, 200                       # core.py:460-463
        self._tkin = None
        self._dens = None          # dict canonical-upper -> array
        self._cdmol = None
        self._abundance = None
        self._use_thermal_opr = False
        self._results = None
        self.escapeProbGeom = escapeProbGeom
        self.deltav = deltav

        if temperature is None:
            raise TypeError("Must specify tkin")
        self._tkin = self._check_temperature(temperature)
        dens_in = collider_densities if collider_densities is not None else (
            total_density if total_density is not None else density)
        if dens_in is not None:
            self.density = dens_in
        col_in = column_per_bin if column_per_bin is not None else column
        if col_in is not None:
            self.column_per_bin = col_in
        if abundance is not None:
            self.abundance = abundance
        self._validate_colliders()
        self.tbg = tbackground
        self.source_area = source_area

    # ---- species / data file (base_class.py:117-139) ------------------------------------------
    @property
    def species(self):
        return self._species

    @species.setter
    def species(self, species):
        if self._species == species:
            return
        molpath = species if os.path.isfile(species) else os.path.join(self.datapath, species + ".dat")
        if not os.path.exists(molpath):
            raise ValueError("Must specify a valid path to a molecular data file else RADEX will crash."
                             "  Current path is {0}".format(molpath))
        self._species = species
        self.molpath = molpath
        self.molfile_is_synthetic = is_synthetic_table(molpath)
        self.mol = get_moldata(molpath)        # host parse only; the GPU context is created on first use
        ids = {v: k for k, v in _COLLIDER_IDS.items()}
        self._valid_colliders = [_CANON[ids[int(i)]] for i in self.mol.partner_id]
        self._results = None

    @property
    def _ctx(self):
        return get_context(self.molpath, self.device)

    @property
    def valid_colliders(self):
        return self._valid_colliders

    # ---- set_params (core.py:388-438) ------------------------------------------------------------
    def set_params(self, density=None, collider_densities=None, column=None, column_per_bin=None,
                   temperature=None, abundance=None, species=None, deltav=None, tbg=None, escapeProbGeom=None):
        if species is not None:
            self.species = species
        if deltav is not None:
            self.deltav = deltav
        if temperature is not None:       # before density, so a thermal OPR sees the new T (core.py:399-402)
            self._tkin = self._check_temperature(temperature)
        if collider_densities is not None:
            self.density = collider_densities
        elif density is not None:
            self.density = density
        if column is not None:
            self.column = column
        elif column_per_bin is not None:
            self.column_per_bin = column_per_bin
        if temperature is not None:
            self.temperature = temperature
        if abundance is not None:
            self.abundance = abundance
        if tbg is not None:
            self.tbg = tbg
        if escapeProbGeom is not None:
            self.escapeProbGeom = escapeProbGeom

    # ---- density (core.py:473-579) ---------------------------------------------------------------
    @property
    def density(self):
        """dict collider -> cm^-3, as the reference's ImmutableDict view of cphys.density."""
        d = {name: np.zeros(1) for name in _CANON.values()}
        if self._dens:
            for k, v in self._dens.items():
                d[_CANON[k]] = v
        return {k: (v if v.size > 1 else float(v[0])) for k, v in d.items()}

    @density.setter
    def density(self, collider_density):
        self._use_thermal_opr = False
        if not isinstance(collider_density, dict):
            collider_density = {"H2": collider_density}        # "Assuming the density is n(H_2)." core.py:503-506
        cd = {}
        for k, v in collider_density.items():
            ku = k.upper()
            if ku not in _COLLIDER_IDS:
                raise ValueError("Collider %s is not one of the valid colliders: %s" % (k, _CANON))
            cd[ku] = np.atleast_1d(np.asarray(_unitless(v), dtype=np.float64))
        prev = dict(self._dens) if self._dens else {}
        new = {k: prev.get(k, np.zeros(1)) for k in ("PH2", "OH2")}
        has_op = any(k in cd and np.any(cd[k] != 0) for k in ("OH2", "PH2"))
        if has_op:
            for k in ("PH2", "OH2"):
                if k in cd:
                    new[k] = cd[k]
        elif "H2" in cd:
            warnings.warn("Using a default ortho-to-para ratio (which will only affect species for which "
                          "independent ortho & para collision rates are given)")
            self._use_thermal_opr = True
            T = np.atleast_1d(self._tkin)
            with np.errstate(all="ignore"):
                opr = np.where(T > 0, np.minimum(3.0, 9.0 * np.exp(-170.6 / T)), 3.0)   # core.py:537-546
            fortho = opr / (1 + opr)
            new["PH2"] = cd["H2"] * (1 - fortho)
            new["OH2"] = cd["H2"] * fortho
        vc = [x.lower() for x in self.valid_colliders]
        if "h2" in vc:                                          # core.py:551-556
            new["H2"] = new["PH2"] + new["OH2"]
            new["PH2"] = np.zeros(1)
            new["OH2"] = np.zeros(1)
        else:
            new["H2"] = np.zeros(1)
        for k in ("E", "H", "HE", "H+"):
            new[k] = cd.get(k, np.zeros(1))
        self._dens = new
        self._validate_colliders()
        self._results = None
        if self._abundance is not None and getattr(self, "_locked_parameter", None) == "abundance":
            self._cdmol = self.total_density * PC_CM * self._abundance

    @property
    def total_density(self):
        """cm^-3 (core.py:586-592): sum over all colliders."""
        tot = 0.0
        for v in (self._dens or {}).values():
            tot = tot + v
        tot = np.atleast_1d(tot)
        return tot if tot.size > 1 else float(tot[0])

    @property
    def opr(self):
        return self._dens["PH2"] / self._dens["OH2"]       # sic: core.py:595-596 returns para/ortho

    def _validate_colliders(self):
        """base_class.py:224-263."""
        if self._dens is None:
            return
        valid = [x.lower() for x in self.valid_colliders]
        matched = [c for c in valid if np.any(self._dens.get(c.upper(), np.zeros(1)) > 0)]
        if not matched:
            raise ValueError("The colliders in the data file {0} have density 0.".format(self.molpath))
        bad = []
        for k, v in self._dens.items():
            kl = k.lower()
            if np.any(v > 0) and kl not in valid:
                if kl in ("oh2", "ph2") and "h2" in matched:
                    continue
                if kl == "h2" and ("oh2" in matched or "ph2" in matched):
                    continue
                bad.append(_CANON[k])
        if bad:
            raise ValueError("There are colliders with specified densities >0 that do not have corresponding "
                             "collision rates.  The bad colliders are {0}".format(bad))

    # ---- temperature (core.py:723-753) -----------------------------------------------------------
    @staticmethod
    def _check_temperature(tkin):
        if tkin is None:
            raise TypeError("Must specify tkin")
        t = np.atleast_1d(np.asarray(_unitless(tkin), dtype=np.float64))
        if np.any(~(t > 0)) or np.any(t > 1e4):
            raise ValueError("Must have kinetic temperature > 0 and < 10^4 K")
        return t

    @property
    def temperature(self):
        """K."""
        return self._tkin if self._tkin.size > 1 else float(self._tkin[0])

    @temperature.setter
    def temperature(self, tkin):
        self._tkin = self._check_temperature(tkin)
        self._results = None
        if self._use_thermal_opr:                               # core.py:748-753
            tot = self._dens["H2"] + self._dens["OH2"] + self._dens["PH2"]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.density = {"H2": tot}

    # ---- column (core.py:755-787) ----------------------------------------------------------------
    @property
    def column(self):
        return self.column_per_bin

    @column.setter
    def column(self, value):
        self.column_per_bin = value

    @property
    def column_per_bin(self):
        """cm^-2."""
        return self._cdmol if self._cdmol.size > 1 else float(self._cdmol[0])

    @column_per_bin.setter
    def column_per_bin(self, col):
        c = np.atleast_1d(np.asarray(_unitless(col), dtype=np.float64))
        if np.any(~(c >= 1e5)) or np.any(c > 1e25):
            raise ValueError("Extremely low or extremely high column.")
        self._cdmol = c
        self._locked_parameter = "column"
        self._results = None

    @property
    def abundance(self):
        if self._abundance is not None:
            return self._abundance
        return self._cdmol / (np.atleast_1d(self.total_density) * PC_CM)

    @abundance.setter
    def abundance(self, abund):
        """column = abundance * n_total * 1 pc (core.py:805-817)."""
        self._abundance = np.atleast_1d(np.asarray(abund, dtype=np.float64))
        if self._dens is not None:
            self.column_per_bin = np.atleast_1d(self.total_density) * PC_CM * self._abundance
        self._locked_parameter = "abundance"

    @property
    def locked_parameter(self):
        return getattr(self, "_locked_parameter", "density")

    @property
    def length(self):
        return PC_CM

    # ---- deltav / tbg / geometry (core.py:819-854, 690-700) ----------------------------------------
    @property
    def deltav(self):
        """km/s."""
        return self._deltav

    @deltav.setter
    def deltav(self, dv):
        dv = float(_unitless(dv))
        self._deltav = dv
        self._results = None

    @property
    def tbg(self):
        """K."""
        return self._tbg

    @tbg.setter
    def tbg(self, tbg):
        if tbg is None:
            return
        self._tbg = float(_unitless(tbg))
        self._results = None

    @property
    def escapeProbGeom(self):
        return {2: "lvg", 1: "sphere", 3: "slab"}[self._method]

    @escapeProbGeom.setter
    def escapeProbGeom(self, g):
        if g not in GEOM:
            raise ValueError("Invalid escapeProbGeom, must be one of " + ",".join(GEOM))
        self._method = GEOM[g]
        self._results = None

    # ---- run_radex (core.py:856-925) ---------------------------------------------------------------
    def _batch_inputs(self):
        if self._dens is None or self._cdmol is None:
            raise ValueError("Must specify two of column, density, and abundance.")
        arrays = [self._tkin, self._cdmol] + [self._dens[k] for k in self._dens]
        n = max(a.size for a in arrays)
        for a in arrays:
            if a.size not in (1, n):
                raise ValueError("batched inputs must share one length")
        ids = {v: k for k, v in _COLLIDER_IDS.items()}
        dens = np.empty((n, self.mol.npart), dtype=np.float64)
        for p, pid in enumerate(self.mol.partner_id):
            dens[:, p] = np.broadcast_to(self._dens[ids[int(pid)]], (n,))
        tk = np.ascontiguousarray(np.broadcast_to(self._tkin, (n,)), dtype=np.float64)
        cd = np.ascontiguousarray(np.broadcast_to(self._cdmol, (n,)), dtype=np.float64)
        return n, tk, dens, cd

    def run_radex(self, silent=True, reuse_last=False, reload_molfile=True, abs_convergence_threshold=1e-16,
                  rel_convergence_threshold=1e-8, validate_colliders=True, stop_rule=STOP_PYRADEX):
        """Solve every model of the current batch on the GPU; returns the iteration counter(s)
        exactly as ``Radex.run_radex`` does.  ``rel_convergence_threshold`` is accepted for
        signature compatibility; in the reference it is dead code (0/0 over the padded xpop)."""
        if validate_colliders:
            self._validate_colliders()
        n, tk, dens, cd = self._batch_inputs()
        nl, nn = self.mol.nlev, self.mol.nline
        out = dict(xpop=np.empty((n, nl)), tex=np.empty((n, nn)), tau=np.empty((n, nn)), surf=np.empty((n, nn)),
                   niter=np.empty(n, np.int32), status=np.empty(n, np.int32))
        opts = _lib.default_opts(stop_rule=stop_rule, miniter=self.miniter, maxiter=self.maxiter,
                                 abs_tol=abs_convergence_threshold)
        import ctypes as C
        _lib.check(_lib.load().rb_solve_batch(
            self._ctx.handle, n, _lib.ptr(tk), _lib.ptr(dens), _lib.ptr(cd), self._deltav, self._tbg, self._method,
            C.byref(opts), _lib.ptr(out["xpop"]), _lib.ptr(out["tex"]), _lib.ptr(out["tau"]), _lib.ptr(out["surf"]),
            _lib.ptr(out["niter"]), _lib.ptr(out["status"])))
        self._results = out
        self._n = n
        if not silent:
            for k in range(n):
                if out["status"][k] & _lib.ST_MAXITER:
                    print("Did not converge in %i iterations, stopping." % self.maxiter)
                else:
                    print("Stopped changing after %i iterations" % out["niter"][k])
        self._iter_counter = out["niter"] if n > 1 else int(out["niter"][0])
        return self._iter_counter

    def _res(self, key):
        if self._results is None:
            self.run_radex(validate_colliders=False)
        a = self._results[key]
        return a if self._n > 1 else a[0]

    # ---- result views (core.py:703-721, 927-1003) ---------------------------------------------------
    @property
    def level_population(self):
        return self._res("xpop")

    @property
    def tex(self):
        """K."""
        return self._res("tex")

    Tex = tex

    @property
    def tau(self):
        return self._res("tau")

    @property
    def status(self):
        return self._res("status")

    @property
    def frequency(self):
        """GHz (the file's frequency column, as radi.spfreq)."""
        return self.mol.spfreq

    @property
    def upperlevelindex(self):
        return self.mol.iupp - 1

    @property
    def lowerlevelindex(self):
        return self.mol.ilow - 1

    @property
    def upperlevelpop(self):
        return self.level_population[..., self.upperlevelindex]

    @property
    def lowerlevelpop(self):
        return self.level_population[..., self.lowerlevelindex]

    @property
    def upperstateenergy(self):
        """K."""
        return self.mol.eup

    @property
    def upperlevel_statisticalweight(self):
        return self.mol.gstat[self.upperlevelindex]

    @property
    def lowerlevel_statisticalweight(self):
        return self.mol.gstat[self.lowerlevelindex]

    @property
    def background_brightness(self):
        """erg s^-1 cm^-2 Hz^-1 sr^-1: backrad's Planck function at tbg (radex.so@0x1be30)."""
        xnu = self.mol.xnu
        hnu = FK_RADEX * xnu / self._tbg
        with np.errstate(over="ignore"):
            return np.where(hnu >= 160.0, 1.0e-30, THC_RADEX * (xnu * xnu * xnu) / (np.exp(hnu) - 1.0))

    @property
    def source_line_surfbrightness(self):
        """erg s^-1 cm^-2 Hz^-1 sr^-1 (base_class.py:275-277), computed by the kernel epilogue."""
        return self._res("surf")

    @property
    def source_brightness(self):
        return self.source_line_surfbrightness + self.background_brightness

    @property
    def source_line_brightness_temperature(self):
        """K: I_nu c^2 / (2 k nu^2) at the rest frequency (base_class.py:298-311)."""
        nu = self.mol.spfreq * 1e9
        return self.source_line_surfbrightness * C_CGS ** 2 / (2.0 * KB_CGS * nu * nu)

    T_B = source_line_brightness_temperature

    @property
    def source_area(self):
        return getattr(self, "_source_area", None)

    @source_area.setter
    def source_area(self, v):
        self._source_area = v

    def get_table(self):
        """Per-line table (base_class.py:361-390) as a pandas DataFrame (scalar inputs) or a dict of
        arrays with a leading batch axis."""
        cols = dict(Tex=self.tex, tau=self.tau, frequency=self.frequency, upperstateenergy=self.upperstateenergy,
                    upperlevel=self.upperlevelindex, lowerlevel=self.lowerlevelindex,
                    upperlevelpop=self.upperlevelpop, lowerlevelpop=self.lowerlevelpop,
                    brightness=self.source_line_surfbrightness, T_B=self.T_B)
        if self._n == 1:
            import pandas as pd
            return pd.DataFrame(cols)
        return cols
