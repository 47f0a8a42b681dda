"""Flat LambdaCDM angular-diameter distance without astropy.

The drivers use ``FlatLambdaCDM(H0=67.8, Om0=0.308)`` (emcee/emcee_radex.py:93) only for
``R_angle`` (emcee/emcee_radex.py:422).  astropy's default Tcmb0=0 means no radiation term.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import quad

C_KMS = 299792.458
H0 = 67.8
OM0 = 0.308


def angular_diameter_distance_mpc(z: float, h0: float = H0, om0: float = OM0) -> float:
    inv_e = lambda x: 1.0 / np.sqrt(om0 * (1.0 + x) ** 3 + (1.0 - om0))
    dc, _ = quad(inv_e, 0.0, float(z), epsabs=0.0, epsrel=1e-12)
    return C_KMS / h0 * dc / (1.0 + float(z))


def r_angle(z: float) -> float:
    """Solid angle [sr] of a 7 kpc source magnified x10 (emcee/emcee_radex.py:421-422)."""
    return ((7.0 / (angular_diameter_distance_mpc(z) * 1000.0)) ** 2 * np.pi) * 10.0
