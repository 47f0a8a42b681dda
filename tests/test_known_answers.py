"""The reference's own known-answer tests (emcee/pyradex/tests/test_radex.py:99-115 `test_radex_results`, :175-200
`test_mod_params`), mirrored on the new `Radex` class.  Their expected numbers belong to the real LAMDA `co.dat`, which
neither the reference tree nor this repository ships: the tests run when RADEX_DATAPATH (or radex_moldata/) holds a real
file and skip otherwise -- with the synthetic table of this repository the numbers are different by construction."""
import os
import warnings

import numpy as np
import pytest

from radex_emcee_b200.radex import Radex, is_synthetic_table

pytestmark = pytest.mark.gpu


def real_co_datapath():
    for d in (os.environ.get("RADEX_DATAPATH"), "radex_moldata", "examples"):
        if d and os.path.exists(os.path.join(d, "co.dat")) and not is_synthetic_table(os.path.join(d, "co.dat")):
            return d
    return None


needs_real_file = pytest.mark.skipif(real_co_datapath() is None,
                                     reason="needs the real LAMDA co.dat (set RADEX_DATAPATH); only the synthetic table is here")


@needs_real_file
def test_radex_results():
    """test_radex.py:99-115: CO, n(H2) = 1e4 (thermal OPR), N = 1e14, dv = 1 km/s, T = 30 K, tbg = 2.73 K, LVG."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rdx = Radex(species="co", datapath=real_co_datapath(), collider_densities={"H2": 1e4}, column_per_bin=1e14,
                    deltav=1.0, temperature=30, tbackground=2.73)
    rdx.run_radex()
    assert rdx.temperature == 30.0 and rdx.column == 1e14
    np.testing.assert_approx_equal(rdx.tex[0], 56.131, 5)
    np.testing.assert_approx_equal(rdx.tau[0], 1.786e-3, 4)
    np.testing.assert_approx_equal(rdx.upperlevelpop[0], 3.640e-1, 4)
    np.testing.assert_approx_equal(rdx.lowerlevelpop[0], 1.339e-1, 4)


@needs_real_file
def test_mod_params():
    """test_radex.py:175-200: parameters changed one at a time on a live object."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        RR = Radex(datapath=real_co_datapath(), species="co", column=1e15, density=1e3, temperature=20)
        tbl = RR()
        np.testing.assert_almost_equal(tbl["Tex"][0], 8.69274406690759, decimal=2)
        RR.column = 1e14
        tbl = RR()
        np.testing.assert_almost_equal(tbl["Tex"][0], 8.0986662583317646, decimal=2)
        RR.density = 1e4
        tbl = RR()
        np.testing.assert_almost_equal(tbl["Tex"][0], 25.381267019506591, decimal=1)
        RR.temperature = 25
        tbl = RR()
        np.testing.assert_almost_equal(tbl["Tex"][0], 37.88, decimal=1)
        RR.deltav = 5
        np.testing.assert_almost_equal(RR.deltav, 5)
        tbl = RR()
        np.testing.assert_almost_equal(tbl["Tex"][0], 37.83, decimal=1)


def test_the_same_sequences_run_on_the_synthetic_table(oracle):
    """The two call sequences above on the table this repository ships: every step against the oracle (1e-5), and the
    qualitative behaviour the reference's numbers show (Tex(1-0) rises with density, is super-thermal at n = 1e4)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rdx = Radex(species="co", collider_densities={"H2": 1e4}, column_per_bin=1e14, deltav=1.0, temperature=30,
                    tbackground=2.73)
        rdx.run_radex()
        opr = min(3.0, 9.0 * np.exp(-170.6 / 30.0))
        fo = opr / (1 + opr)
        ref = oracle.solve_batch([30.0], [1e4 * (1 - fo)], [1e4 * fo], [1e14], tbg=2.73)
        np.testing.assert_allclose(rdx.tex[:6], ref["tex"][0][:6], rtol=1e-5)
        np.testing.assert_allclose(rdx.tau[:6], ref["tau"][0][:6], rtol=1e-5)
        RR = Radex(species="co", column=1e15, density=1e3, temperature=20)
        seq = []
        for change in (None, ("column", 1e14), ("density", 1e4), ("temperature", 25), ("deltav", 5)):
            if change:
                setattr(RR, *change)
            tbl = RR()
            seq.append(float(tbl["Tex"][0]))
            T, n, N, dv = RR.temperature, 1e3 if len(seq) < 3 else 1e4, RR.column, RR.deltav
            o = min(3.0, 9.0 * np.exp(-170.6 / T))
            r = oracle.solve_batch([T], [n / (1 + o)], [n * o / (1 + o)], [N], deltav_kms=dv, tbg=2.7315)
            assert abs(seq[-1] / r["tex"][0][0] - 1) < 1e-5, (change, seq[-1], r["tex"][0][0])
    assert seq[2] > seq[1] and len(set(seq)) == 5
