"""Host logic of the per-source pipeline (driver.py): finite-difference batching, summaries; and, on a GPU,
one short end-to-end fit (reference: main() of emcee/emcee_radex.py:382-531, emcee_radex_2comp.py:479-608)."""
import os
import pickle

import numpy as np
import pytest

from radex_emcee_b200 import driver
from radex_emcee_b200.data import DATA_DIR, get_source, read_data


def test_forward_steps_follow_scipy_two_point_rule():
    from scipy.optimize._numdiff import approx_derivative
    lo, hi = np.array([0.0, -5.0, 1.0]), np.array([2.0, 5.0, 3.0])
    x = np.array([2.0, -1.5, 1.0])                # first at its upper bound, last at its lower bound
    h = driver.forward_steps(x, lo, hi)
    assert h[0] < 0 and h[1] < 0 and h[2] > 0     # flipped at the upper bound; sign(x) elsewhere
    assert np.all((x + h >= lo) & (x + h <= hi))
    f = lambda p: np.array([p[0] ** 2 + p[1], np.sin(p[2]) * p[0], p[1] * p[2]])
    fb = lambda P: np.array([f(p) for p in np.atleast_2d(P)])
    J = driver.fd_jacobian(fb, x, lo, hi)
    Jref = approx_derivative(f, x, method="2-point", bounds=(lo, hi))
    np.testing.assert_allclose(J, Jref, rtol=0, atol=1e-12)


def test_value_and_gradient_is_one_batched_call():
    calls = []

    def fun(P):
        calls.append(np.array(P))
        return np.sum(np.asarray(P) ** 2, axis=1)

    x = np.array([1.0, -2.0, 0.5, 3.0])
    f, g = driver.fd_value_and_gradient(fun, x, np.full(4, -10.0), np.array([10.0, 10.0, 10.0, 3.0]))
    assert len(calls) == 1 and calls[0].shape == (5, 4)
    assert f == pytest.approx(14.25)
    np.testing.assert_allclose(g, 2 * x, atol=1e-6)
    assert calls[0][4, 3] < 3.0                  # at the upper bound the step goes inwards


def test_summaries():
    rng = np.random.default_rng(0)
    chain = rng.normal([4.0, 1.5, 17.0, -10.0, 3.0, 2.5, 16.0, -11.0], 0.1, size=(20000, 8))
    s = driver.posterior_summary(chain, 2)
    assert len(s) == 2 and set(s[0]) == {"n_H2", "T_kin", "N_CO/dv", "P"}
    med, up, dn = s[0]["P"]
    assert med == pytest.approx(5.5, abs=0.01) and up == pytest.approx(0.1 * np.sqrt(2), rel=0.05)
    assert s[1]["n_H2"][0] == pytest.approx(3.0, abs=0.01)
    X = rng.normal(size=(500, 3)) * [1.0, 10.0, 0.1]
    for metric in ("mahalanobis", "z", "euclidean"):
        x, i, d2 = driver.nearest_sample_to_vector(X, X[17] + 1e-9, metric=metric)
        assert i == 17 and d2 < 1e-10 and np.array_equal(x, X[17])


@pytest.mark.gpu
@pytest.mark.parametrize("ncomp", [1, 2])
def test_fit_source_end_to_end(tmp_path, ncomp):
    data = read_data(os.path.join(DATA_DIR, "flux.dat" if ncomp == 1 else "flux_for2p.dat"))
    with pytest.raises(ValueError, match="synthetic"):       # the shipped table is for tests: refuse unless told
        driver.fit_source("G09v1.97", data, ncomp=ncomp, nwalkers=64, n_iter_burn=1, n_iter_walk=1)
    res = driver.fit_source("G09v1.97", data, ncomp=ncomp, nwalkers=64, n_iter_burn=10, n_iter_walk=20,
                            outdir=str(tmp_path), allow_synthetic=True)
    assert res["molfile_synthetic"] and res["molfile"].endswith("co.dat")
    assert res["acceptance_fraction_per_walker"].shape == (64,)
    b = res["bounds"]
    assert np.all((res["popt"] >= b[:, 0]) & (res["popt"] <= b[:, 1]))
    assert np.all((res["pmin"] >= b[:, 0]) & (res["pmin"] <= b[:, 1]))
    assert res["chain"].shape == (20, 64, 4 * ncomp) and res["lnprobability"].shape == (20, 64)
    assert np.isfinite(res["lnprobability"]).mean() > 0.9 and 0.05 < res["acceptance_fraction"] < 0.95
    # curve_fit improved on the starting point: chi^2 of the SLED at popt is below the one at p0
    mod = driver._modules(ncomp)
    from radex_emcee_b200.radex import Radex
    tbg, _ra, bounds, p0 = mod.source_setup(res["z"])
    R = Radex(species="co", density={"oH2": 7.5e9, "pH2": 2.5e9}, column=1e6, temperature=20.0, tbackground=tbg)
    Jup, flux, eflux = res["data"]
    chi2 = lambda p: float(np.sum(((flux - mod.model_lvg(Jup, p, R)) / eflux) ** 2))
    assert chi2(res["popt"]) <= chi2(p0) + 1e-9
    # the batched gradients cost one launch per iteration, not ndim + 1
    assert res["prefit_info"]["lnprob_launches"] < 400
    with open(res["pickle"], "rb") as f:
        tup = pickle.load(f)
    assert len(tup) == (8 if ncomp == 1 else 9) and tup[0] == "G09v1.97"
    np.testing.assert_array_equal(tup[-1][0], res["chain"])
    import io
    buf = io.StringIO()
    driver.print_summary(res, ncomp, file=buf)
    assert buf.getvalue().count("xxx:") >= 7


@pytest.mark.gpu
def test_prefit_matches_scipy_on_the_oracle_model(oracle):
    """the pre-sampling optimisers decide the walkers' starting ball.  driver.prefit (GPU model,
    batched finite differences with scipy's own step rules) against plain scipy on the CPU oracle's model_lvg / lnprob
    (emcee_radex.py:444-467: curve_fit with bounds -> trf with a 2-point Jacobian; minimize -> L-BFGS-B with its
    numerical gradient), same starting point."""
    from scipy.optimize import curve_fit, minimize
    from radex_emcee_b200 import emcee_radex as er1
    from radex_emcee_b200.radex import Radex
    data = read_data(os.path.join(DATA_DIR, "flux.dat"))
    for name in ("G09v1.97", "NCv1.143"):
        z, lw, Jup, flux, eflux = get_source(name, data)
        tbg, _ra, bounds, p0 = er1.source_setup(z)
        lo, hi = bounds[:, 0], bounds[:, 1]
        R = Radex(species="co", density={"oH2": 7.5e9, "pH2": 2.5e9}, column=1e6, temperature=20.0, tbackground=tbg)
        popt, pcov, pmin, info = driver.prefit(1, R, Jup, flux, eflux, bounds, p0)

        def model_cpu(_x, *p):
            p = np.asarray(p)
            n = 10.0 ** p[0]
            r = oracle.solve_batch([10.0 ** p[1]], [0.25 * n], [0.75 * n], [10.0 ** p[2]], tbg=tbg)
            return r["surf"][0][np.asarray(Jup) - 1] * 10.0 ** p[3] * 1e23

        popt_ref, _ = curve_fit(model_cpu, Jup, flux, sigma=eflux, p0=p0, bounds=(lo, hi))
        chi2 = lambda p: float(np.sum(((flux - model_cpu(None, *p)) / eflux) ** 2))
        # same minimum: chi^2 to 1e-3 relative and the parameters to 0.01 dex along well-constrained directions
        assert chi2(popt) == pytest.approx(chi2(popt_ref), rel=1e-3, abs=1e-6), (name, popt, popt_ref)
        assert np.abs(popt - popt_ref).max() < 0.05, (name, popt, popt_ref)
        nll = lambda p: -oracle.lnprob1(p, Jup, flux, eflux, bounds, tbg)
        res = minimize(nll, popt_ref, bounds=list(zip(lo, hi)), method="L-BFGS-B")
        assert nll(pmin) == pytest.approx(res.fun, rel=1e-3, abs=1e-4), (name, pmin, res.x)
        assert np.abs(pmin - res.x).max() < 0.05, (name, pmin, res.x)


@pytest.mark.gpu
def test_sources_fitted_concurrently_through_the_driver(tmp_path):
    """BASELINE.json configs[3] at the driver level: three sources, one ensemble, one pickle per source in the
    reference's layout; every source's chain stays inside its own bounds and finds a finite posterior."""
    data = read_data(os.path.join(DATA_DIR, "flux.dat"))
    names = list(data)[:3]
    out = driver.fit_sources_concurrently(names, data, ncomp=1, nwalkers=40, n_iter_burn=10, n_iter_walk=15,
                                          outdir=str(tmp_path), allow_synthetic=True)
    assert [r["source"] for r in out] == names
    for r in out:
        assert r["chain"].shape == (15, 40, 4) and r["lnprobability"].shape == (15, 40)
        b = r["bounds"]
        assert np.all((r["chain"] >= b[:, 0]) & (r["chain"] <= b[:, 1]))
        assert np.isfinite(r["lnprobability"]).mean() > 0.9 and 0.05 < r["acceptance_fraction"] < 0.95
        with open(r["pickle"], "rb") as f:
            tup = pickle.load(f)
        assert len(tup) == 8 and tup[0] == r["source"]
        np.testing.assert_array_equal(tup[-1][0], r["chain"])
    assert len({tuple(r["bounds"][1]) for r in out}) == 3          # every source has its own background temperature
