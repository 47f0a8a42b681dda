"""CPU tests of the host-side mirror of the reference interface: flux-table readers, cosmology,
priors, the Radex parameter protocol and its ValueErrors (no kernel launches)."""
import warnings

import numpy as np
import pytest

from conftest import ROOT
from radex_emcee_b200 import emcee_radex as er1
from radex_emcee_b200 import emcee_radex_2comp as er2
from radex_emcee_b200.cosmo import angular_diameter_distance_mpc, r_angle
from radex_emcee_b200.data import get_source, read_data
from radex_emcee_b200.radex import Radex


def test_read_data_one_component():
    d = read_data(ROOT + "/data/flux.dat")
    assert len(d) == 16 and list(d)[0] == "G09v1.97" and "NAv1.195" in d
    z, lw, jup, flux, eflux = get_source("G09v1.97", d)
    assert z == 3.6345 and lw == 348.3
    np.testing.assert_array_equal(jup, [3, 4, 5, 6, 7])
    np.testing.assert_array_equal(flux, [5.699, 7.800, 9.734, 9.979, 7.962])
    np.testing.assert_array_equal(eflux, [2.248, 1.500, 1.188, 1.672, 0.915])
    # per-source line sets of SURVEY.md appendix B
    expect = {"G09v1.40": [2, 4, 6, 7], "SDP81": [1, 3, 5, 8, 10], "G12v2.30": [1, 3, 4, 5, 6, 8, 11],
              "NAv1.195": [5], "G15v2.779": [4, 5, 7]}
    for k, v in expect.items():
        assert list(get_source(k, d)[2]) == v


def test_read_data_two_component():
    d = read_data(ROOT + "/data/flux_for2p.dat")
    assert len(d) == 15 and "NAv1.195" not in d          # commented row dropped
    z, T_d, lw, jup, flux, eflux = get_source("G09v1.97", d)
    assert (z, T_d, lw) == (3.6345, 44.0, 348.3) and list(jup) == [3, 4, 5, 6, 7]


def test_read_data_errors(tmp_path):
    bad = tmp_path / "bad.dat"
    bad.write_text("# c\nA 1 2 3 4 5 6 7 8 9 10\nB 1 2 3\n")
    with pytest.raises(ValueError, match="columns"):
        read_data(str(bad))


def test_cosmology():
    # flat LCDM H0=67.8 Om0=0.308 (emcee_radex.py:93): D_A = D_L/(1+z)^2 with the D_L column of flux.dat
    for z, dl in ((3.6345, 32751.0), (2.0924, 16835.0), (2.3053, 18942.0), (4.243, 39411.0)):
        da = angular_diameter_distance_mpc(z)
        assert abs(da * (1 + z) ** 2 / dl - 1) < 2e-3
    assert abs(np.log10(r_angle(3.6345)) + 9.179) < 0.01
    tbg, ra, b, p0 = er1.source_setup(3.6345)
    assert abs(tbg - 12.65913675) < 1e-12 and b.shape == (4, 2) and abs(b[3, 1] - b[3, 0] - 8) < 1e-12
    np.testing.assert_array_equal(p0, [4.0, 1.4, 17.8, -9.85])
    tbg, ra, b, p0 = er2.source_setup(3.6345)
    assert b.shape == (8, 2) and abs(b[3, 1] - b[3, 0] - 18) < 1e-12


def test_lnprior_matches_oracle(oracle):
    L = oracle.L
    rng = np.random.default_rng(0)
    tbg, ra, b1, p0 = er1.source_setup(3.6345)
    P = rng.uniform(b1[:, 0] - 0.3, b1[:, 1] + 0.3, size=(4000, 4))
    got = er1.lnprior(P, b1)
    ref = np.array([L.ro_lnprior1(p.ctypes.data_as(L.ro_lnprior1.argtypes[0]),
                                  np.ascontiguousarray(b1).ctypes.data_as(L.ro_lnprior1.argtypes[1])) for p in P])
    np.testing.assert_array_equal(got, ref)
    assert 0.05 < np.isfinite(got).mean() < 0.9
    tbg, ra, b2, p0 = er2.source_setup(3.6345)
    P = p0 + rng.standard_normal((4000, 8)) * 0.8
    for td in (44.0, None, -1.0):
        got = er2.lnprior(P, b2, T_d=td)
        b2c = np.ascontiguousarray(b2)
        ref = np.array([L.ro_lnprior2(np.ascontiguousarray(p).ctypes.data_as(L.ro_lnprior2.argtypes[0]),
                                      b2c.ctypes.data_as(L.ro_lnprior2.argtypes[1]), int(td is not None),
                                      float(td or 0.0)) for p in P])
        assert ((got == -np.inf) == (ref == -np.inf)).all()
        fin = np.isfinite(ref)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-13)
    # the flat terms add minus the WIDTH of each bound (emcee_radex_2comp.py:233), T_d term Gaussian with sigma=T_d
    p = p0.copy()
    val = er2.lnprior(p, b2, T_d=None)
    assert abs(val + (b2[:, 1] - b2[:, 0]).sum()) < 1e-12


def test_radex_parameter_protocol():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R = Radex(species="co", collider_densities={"H2": 1e3}, column_per_bin=1e13, temperature=20)
    assert abs(R.total_density - 1e3) < 1e-9                      # test_selfconsistent_density
    opr = min(3.0, 9.0 * np.exp(-170.6 / 20.0))
    assert abs(R.density["oH2"] / R.density["pH2"] - opr) < 1e-12   # thermal OPR (core.py:537-546)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R.temperature = 30
    assert abs(R.total_density - 1e3) < 1e-9
    opr = min(3.0, 9.0 * np.exp(-170.6 / 30.0))
    assert abs(R.density["oH2"] / R.density["pH2"] - opr) < 1e-12
    R.density = {"oH2": 990, "pH2": 10}
    assert abs(R.total_density - 1e3) < 1e-9 and R.density["oH2"] == 990
    R.density = {"ph2": 20}                                        # case-insensitive; oH2 persists (core.py:525-528)
    assert R.density["oH2"] == 990 and R.density["pH2"] == 20
    assert R.valid_colliders == ["pH2", "oH2"] and R.escapeProbGeom == "lvg"
    assert R.locked_parameter == "column"
    R.abundance = 1e-9
    assert R.locked_parameter == "abundance"
    np.testing.assert_allclose(R.column, 1e-9 * 1010 * 3.0856775814913673e18)


def test_radex_errors():
    kw = dict(species="co", density={"oH2": 750.0, "pH2": 250.0}, column=1e15, temperature=20.0)
    with pytest.raises(ValueError, match="two of column"):
        Radex(species="co", abundance=1e-4, column=1e15, density=1e3)
    with pytest.raises(ValueError, match="Must specify two"):
        Radex(species="co", temperature=20, column=None)
    with pytest.raises(ValueError, match="one of density"):
        Radex(species="co", density=1e3, total_density=1e3, column=1e15, temperature=20)
    with pytest.raises(ValueError, match="valid path"):
        Radex(species="nosuchmolecule", density={"oH2": 1.0}, column=1e15, temperature=20)
    with pytest.raises(ValueError, match="kinetic temperature"):
        Radex(**dict(kw, temperature=2e4))
    with pytest.raises(ValueError, match="column"):
        Radex(**dict(kw, column=1e3))
    with pytest.raises(ValueError, match="escapeProbGeom"):
        Radex(**dict(kw, escapeProbGeom="cube"))
    with pytest.raises(ValueError, match="density 0"):
        Radex(**dict(kw, density={"oH2": 0.0, "pH2": 0.0}))
    with pytest.raises(ValueError, match="corresponding collision rates"):
        Radex(**dict(kw, density={"oH2": 1.0, "e": 5.0}))
    R = Radex(**kw)
    with pytest.raises(ValueError):
        R.set_params(temperature=0.0)
    R.set_params(tbg=10.0, deltav=5.0, escapeProbGeom="sphere")
    assert (R.tbg, R.deltav, R.escapeProbGeom) == (10.0, 5.0, "sphere")
    bb = R.background_brightness
    assert bb.shape == (40,) and (bb > 0).all()


def test_synthetic_table_is_recognised_and_the_driver_refuses_it(tmp_path):
    """the shipped co.dat is a synthetic table; nothing may fit real data with it silently."""
    from radex_emcee_b200.radex import is_synthetic_table
    from radex_emcee_b200.synth_lamda import default_path, rotor_path
    assert is_synthetic_table(default_path()) and is_synthetic_table(rotor_path())
    real = tmp_path / "co.dat"
    real.write_text("!MOLECULE\nCO\n!MOLECULAR WEIGHT\n28.0\n")
    assert not is_synthetic_table(str(real)) and not is_synthetic_table(str(tmp_path / "missing.dat"))
    from radex_emcee_b200 import driver
    import inspect
    assert "allow_synthetic" in inspect.signature(driver.fit_source).parameters
    assert "datapath" in inspect.signature(driver.fit_sources_concurrently).parameters


def test_walkers_independent_is_emcee_s_check():
    from radex_emcee_b200.sampler import walkers_independent
    rng = np.random.default_rng(0)
    good = rng.standard_normal((40, 4)) * [1e-3, 1.0, 1e3, 1e-6] + [4.0, 1.4, 17.8, -9.85]
    assert walkers_independent(good)                        # scale does not matter: columns are normalised
    dep = good.copy()
    dep[:, 3] = 2.0 * dep[:, 1] - 0.5 * dep[:, 0]
    assert not walkers_independent(dep)
    const = good.copy()
    const[:, 2] = 16.0          # a constant whose mean is exact (emcee compares the centred column with 0)
    assert not walkers_independent(const)
    nan = good.copy()
    nan[3, 1] = np.nan
    assert not walkers_independent(nan)
