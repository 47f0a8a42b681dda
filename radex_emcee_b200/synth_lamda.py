"""Deterministic synthetic CO-like molecular data file in the LAMDA text format.

The reference expects ``radex_moldata/co.dat`` (emcee/emcee_radex.py:110) but does not ship it,
and there is no network here.  This module writes a *CO-like* file with the public LAMDA layout
(SURVEY.md Appendix A): 41 rotational levels of a centrifugally distorted rigid rotor with CO's
spectroscopic constants, 40 dipole lines with A from the CO dipole moment, and two collision
partners (2 = p-H2, 3 = o-H2) x 820 downward rates x 25 temperatures built from smooth,
made-up rate surfaces of realistic magnitude.  Level energies and Einstein A values agree with
the public CO file to ~4 digits; the collision rates are NOT the published ones, so absolute
answers differ from a run with the real co.dat (the loader accepts the real file unchanged).

All numbers are printed with the few significant digits a LAMDA file carries, so that parsing the
text is the single source of truth (no hidden double-precision state).
"""
from __future__ import annotations

import math
import os

# CO spectroscopic constants (cm^-1) and dipole moment (Debye)
_B = 1.922528960
_D = 6.12107e-6
_H = 5.7e-12
_MU_DEBYE = 0.11011
_CLIGHT = 2.99792458e10
_HPLANCK = 6.6260755e-27
_KB = 1.380658e-16

COLL_TEMPS = [2.0, 3.0, 5.0, 7.0, 10.0, 15.0, 20.0, 30.0, 40.0, 50.0, 60.0, 70.0, 80.0, 90.0,
              100.0, 150.0, 200.0, 300.0, 400.0, 500.0, 600.0, 800.0, 1000.0, 2000.0, 3000.0]


def level_energy(j: int) -> float:
    x = j * (j + 1.0)
    return _B * x - _D * x * x + _H * x * x * x


def einstein_a(jup: int, nu_cm: float) -> float:
    """A(J->J-1) = 64 pi^4 nu^3 mu^2 / (3 h c^3) * J/(2J+1), nu in Hz."""
    nu = nu_cm * _CLIGHT
    mu = _MU_DEBYE * 1e-18
    return 64.0 * math.pi ** 4 * nu ** 3 * mu * mu / (3.0 * _HPLANCK * _CLIGHT ** 3) * jup / (2.0 * jup + 1.0)


def _rate(partner: int, jup: int, jlo: int, t: float) -> float:
    """Smooth made-up downward rate coefficient surface (cm^3 s^-1)."""
    dj = jup - jlo
    # base magnitude and a mild propensity for even dJ with p-H2, odd dJ with o-H2
    if partner == 2:
        base = 3.2e-11 * (1.25 if dj % 2 == 0 else 1.0)
        slope = 0.58
    else:
        base = 4.1e-11 * (1.0 if dj % 2 == 0 else 1.3)
        slope = 0.52
    # fall-off with dJ that softens with temperature
    soft = 1.0 + (t / 180.0) ** 0.65
    fall = math.exp(-slope * (dj - 1) * 2.2 / soft)
    # temperature dependence: shallow rise, a low-T resonance-like bump for low J
    trise = (1.0 + t / 60.0) ** 0.32
    bump = 1.0 + 0.35 * math.exp(-((math.log10(t) - 1.1) ** 2) / 0.18) / (1.0 + 0.15 * jlo)
    jdep = 1.0 / (1.0 + 0.012 * jup) * (1.0 + 0.25 * math.exp(-0.5 * jlo))
    return base * fall * trise * bump * jdep


def write_co_synth(path: str, nlev: int = 41) -> str:
    """Write the synthetic file to ``path`` and return the path."""
    lines = []
    w = lines.append
    w("!MOLECULE")
    w("CO (synthetic CO-like table, radex_emcee_b200.synth_lamda)")
    w("!MOLECULAR WEIGHT")
    w("28.0")
    w("!NUMBER OF ENERGY LEVELS")
    w(str(nlev))
    w("!LEVEL + ENERGIES(cm^-1) + WEIGHT + J")
    energies = []
    for j in range(nlev):
        e = float("%.9f" % level_energy(j))
        energies.append(e)
        w("%5d %15.9f %6.1f %5d" % (j + 1, e, 2.0 * j + 1.0, j))
    w("!NUMBER OF RADIATIVE TRANSITIONS")
    w(str(nlev - 1))
    w("!TRANS + UP + LOW + EINSTEINA(s^-1) + FREQ(GHz) + E_u(K)")
    for j in range(1, nlev):
        nu_cm = energies[j] - energies[j - 1]
        a = einstein_a(j, nu_cm)
        w("%5d %5d %5d %11.3e %16.7f %10.2f" % (j, j + 1, j, a, nu_cm * _CLIGHT * 1e-9,
                                                 energies[j] * _HPLANCK * _CLIGHT / _KB))
    w("!NUMBER OF COLL PARTNERS")
    w("2")
    for partner, label in ((2, "2 CO-pH2 synthetic smooth rate surface"),
                           (3, "3 CO-oH2 synthetic smooth rate surface")):
        w("!COLLISIONS BETWEEN")
        w(label)
        w("!NUMBER OF COLL TRANS")
        w(str(nlev * (nlev - 1) // 2))
        w("!NUMBER OF COLL TEMPS")
        w(str(len(COLL_TEMPS)))
        w("!COLL TEMPS")
        w(" ".join("%7.1f" % t for t in COLL_TEMPS))
        w("!TRANS + UP + LOW + COLLRATES(cm^3 s^-1)")
        k = 0
        for jup in range(1, nlev):
            for jlo in range(jup):
                k += 1
                w("%5d %5d %5d " % (k, jup + 1, jlo + 1) +
                  " ".join("%10.3e" % _rate(partner, jup, jlo, t) for t in COLL_TEMPS))
    w("!NOTES: synthetic table; collision rates are not published values")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


# ---- a second, differently shaped table for the general-molecule path -------------------------------------------
# 21 levels of a heavier linear rotor ("CS-like": B = 0.8171 cm^-1, mu = 1.958 D), 20 lines, and two partners that
# exercise what CO's table does not: H2 itself (LAMDA id 1: pyradex folds oH2 + pH2 into it, core.py:551-556) on a
# 9-node temperature grid with every downward pair, and electrons (id 4) on a 6-node grid with |dJ| <= 3 only.
ROTOR_B, ROTOR_D, ROTOR_MU = 0.8171, 1.43e-6, 1.958
ROTOR_TEMPS_H2 = [10.0, 20.0, 40.0, 60.0, 100.0, 150.0, 200.0, 300.0, 500.0]
ROTOR_TEMPS_E = [10.0, 30.0, 100.0, 300.0, 1000.0, 3000.0]


def _rotor_rate(partner: int, jup: int, jlo: int, t: float) -> float:
    dj = jup - jlo
    if partner == 1:
        return 6.5e-11 * math.exp(-0.45 * (dj - 1)) * (1.0 + t / 90.0) ** 0.25 / (1.0 + 0.02 * jup) * \
            (1.15 if dj % 2 == 0 else 1.0)
    # electrons: dipole-dominated, ~1e-6 cm^3/s, falling with dJ and slowly with T
    return 2.4e-6 * (0.08 ** (dj - 1)) * (t / 100.0) ** -0.35 * (jup / (2.0 * jup + 1.0)) * (1.0 + 0.1 * math.exp(-jlo))


def write_rotor_synth(path: str, nlev: int = 21) -> str:
    """Write the 21-level synthetic rotor (partners H2 and e) to ``path``."""
    lines = []
    w = lines.append
    w("!MOLECULE")
    w("ROTOR21 (synthetic CS-like table, radex_emcee_b200.synth_lamda)")
    w("!MOLECULAR WEIGHT")
    w("44.0")
    w("!NUMBER OF ENERGY LEVELS")
    w(str(nlev))
    w("!LEVEL + ENERGIES(cm^-1) + WEIGHT + J")
    energies = []
    for j in range(nlev):
        x = j * (j + 1.0)
        e = float("%.9f" % (ROTOR_B * x - ROTOR_D * x * x))
        energies.append(e)
        w("%5d %15.9f %6.1f %5d" % (j + 1, e, 2.0 * j + 1.0, j))
    w("!NUMBER OF RADIATIVE TRANSITIONS")
    w(str(nlev - 1))
    w("!TRANS + UP + LOW + EINSTEINA(s^-1) + FREQ(GHz) + E_u(K)")
    for j in range(1, nlev):
        nu_cm = energies[j] - energies[j - 1]
        nu = nu_cm * _CLIGHT
        mu = ROTOR_MU * 1e-18
        a = 64.0 * math.pi ** 4 * nu ** 3 * mu * mu / (3.0 * _HPLANCK * _CLIGHT ** 3) * j / (2.0 * j + 1.0)
        w("%5d %5d %5d %11.3e %16.7f %10.2f" % (j, j + 1, j, a, nu * 1e-9, energies[j] * _HPLANCK * _CLIGHT / _KB))
    w("!NUMBER OF COLL PARTNERS")
    w("2")
    for partner, label, temps, maxdj in ((1, "1 ROTOR-H2 synthetic smooth rate surface", ROTOR_TEMPS_H2, nlev),
                                         (4, "4 ROTOR-e synthetic dipole-like rates, |dJ| <= 3", ROTOR_TEMPS_E, 3)):
        pairs = [(ju, jl) for ju in range(1, nlev) for jl in range(ju) if ju - jl <= maxdj]
        w("!COLLISIONS BETWEEN")
        w(label)
        w("!NUMBER OF COLL TRANS")
        w(str(len(pairs)))
        w("!NUMBER OF COLL TEMPS")
        w(str(len(temps)))
        w("!COLL TEMPS")
        w(" ".join("%7.1f" % t for t in temps))
        w("!TRANS + UP + LOW + COLLRATES(cm^3 s^-1)")
        for k, (ju, jl) in enumerate(pairs):
            w("%5d %5d %5d " % (k + 1, ju + 1, jl + 1) + " ".join("%10.3e" % _rotor_rate(partner, ju, jl, t) for t in temps))
    w("!NOTES: synthetic table; collision rates are not published values")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def rotor_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "rotor21.dat")


def default_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "co.dat")


if __name__ == "__main__":
    print(write_co_synth(default_path()))
    print(write_rotor_synth(rotor_path()))
