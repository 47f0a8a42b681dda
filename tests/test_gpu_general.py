"""The general-molecule path (reference: emcee/pyradex/core.py:465-471,492-512,690-700,
base_class.py:224-263): any LAMDA file with up to 64 levels and any mix of the seven collision partners runs through
k_lvg_solve_v1 / k_lnprob_v1 (shared-memory matrix, partial-pivot LU on the reduced system the reference's lubksb
solves).  Checked against the reference binary's fixtures for a second molecule, against the oracle on random sweeps of
that molecule with H2 + electron densities in all three geometries, and on CO itself (kernel=1 against the oracle and
against the default GTH kernels)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, MOLFILE, ROOT, draw_params
from oracle import parity
from radex_emcee_b200 import _lib
from radex_emcee_b200.radex import Radex

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ROTOR = os.path.join(ROOT, "radex_emcee_b200", "data", "rotor21.dat")


@pytest.fixture(scope="module")
def rotor_ctx():
    return _lib.Context(_lib.MolData(ROTOR), 0)


@pytest.fixture(scope="module")
def co_ctx():
    return _lib.Context(_lib.MolData(MOLFILE), 0)


def gpu_solve_dens(ctx, T, dens7, N, tbg, method=2, **optkw):
    """dens7[n, 7] by LAMDA partner id -> the file's partner order, as rb_solve_batch wants it."""
    T, N = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (T, N))
    n, mol = T.size, ctx.mol
    dens = np.ascontiguousarray(np.asarray(dens7, dtype=np.float64)[:, np.asarray(mol.partner_id) - 1])
    out = dict(xpop=np.empty((n, mol.nlev)), tex=np.empty((n, mol.nline)), tau=np.empty((n, mol.nline)),
               surf=np.empty((n, mol.nline)), niter=np.empty(n, np.int32), status=np.empty(n, np.int32))
    opts = _lib.default_opts(**optkw)
    _lib.check(_lib.load().rb_solve_batch(ctx.handle, n, _lib.ptr(T), _lib.ptr(dens), _lib.ptr(N), 1.0, float(tbg), method,
                                          C.byref(opts), _lib.ptr(out["xpop"]), _lib.ptr(out["tex"]), _lib.ptr(out["tau"]),
                                          _lib.ptr(out["surf"]), _lib.ptr(out["niter"]), _lib.ptr(out["status"])))
    return out


def rotor_draws(rng, n, tbg):
    T = 10 ** rng.uniform(np.log10(max(tbg, 5.0)), 2.7, n)
    d = np.zeros((n, 7))
    d[:, 0] = 10 ** rng.uniform(2.5, 7.0, n)
    d[:, 3] = np.where(rng.uniform(size=n) < 0.3, 0.0, 10 ** rng.uniform(-2.0, 2.0, n))
    return T, d, 10 ** rng.uniform(11.5, 15.5, n)


def test_second_molecule_vs_reference_binary_fixtures(rotor_ctx):
    g = np.load(os.path.join(GOLDEN, "macho_solve_rotor21.npz"))
    c = g["cases"]
    for method in (1, 2, 3):
        for tbg in np.unique(c[:, 4]):
            sel = (c[:, 5] == method) & (c[:, 4] == tbg)
            if not sel.any():
                continue
            d = np.zeros((int(sel.sum()), 7))
            d[:, 0], d[:, 3] = c[sel, 1], c[sel, 2]
            got = gpu_solve_dens(rotor_ctx, c[sel, 0], d, c[sel, 3], tbg, method)
            dn = np.abs(got["niter"] - g["niter"][sel])      # the 1e-16 stop test jitters in its last bit
            assert np.median(dn) <= 3 and (dn <= 40).mean() >= 0.8, dn      # one model in eight may end at the cap
            sig = g["xpop"][sel] > 1e-9
            assert (np.abs(got["xpop"] - g["xpop"][sel]) / g["xpop"][sel])[sig].max() < RTOL
            sl = sig[:, rotor_ctx.mol.iupp - 1]
            assert (np.abs(got["tex"] - g["tex"][sel]) / np.abs(g["tex"][sel]))[sl].max() < RTOL
            assert (np.abs(got["tau"] - g["tau"][sel]) / np.maximum(np.abs(g["tau"][sel]), 1e-12))[sl].max() < RTOL


@pytest.mark.parametrize("method,tbg", [(2, 2.7315), (2, 10.926), (1, 2.7315), (3, 2.7315)])
def test_second_molecule_sweep_vs_oracle(rotor_ctx, method, tbg):
    from test_gpu_solve import record
    T, d, N = rotor_draws(np.random.default_rng(40 + method + int(tbg)), 300, tbg)
    got = gpu_solve_dens(rotor_ctx, T, d, N, tbg, method)
    ref, cls, runs = parity.classify(ROTOR, T, d, N, tbg, method)
    w = parity.worst(got, ref, ref["iupp"])
    wp = cls["well_posed"]
    record("rotor21_method%d_tbg%.3f" % (method, tbg), models=300, classes={k: int(v.sum()) for k, v in cls.items()},
           max_err_well_posed=float(w[wp].max()), all_within_tol=int((w < RTOL).sum()))
    assert wp.mean() > 0.9, wp.mean()
    assert w[wp].max() < RTOL, w[wp].max()
    both = wp & (got["niter"] < 200) & (ref["niter"] < 200)
    dn = np.abs(got["niter"] - ref["niter"])[both]
    assert np.median(dn) <= 3 and np.quantile(dn, 0.9) <= 40


def test_co_through_the_general_kernel(co_ctx):
    """kernel=1 on the 41-level table: the pivoted LU of the reference on the GPU, against the oracle and against the
    default kernels (GTH elimination) -- two independent eliminations, one answer."""
    from test_gpu_solve import gpu_solve, record
    P = draw_params(np.random.default_rng(77), 384, 10.926)
    T, nh2, N = P[:, 0], P[:, 1], P[:, 2]
    lu = gpu_solve(co_ctx, T, nh2, N, 10.926, kernel=1)
    gth = gpu_solve(co_ctx, T, nh2, N, 10.926)
    ref, cls, runs = parity.classify(MOLFILE, T, nh2, N, 10.926)
    wp = cls["well_posed"]
    w_lu, w_gth = parity.worst(lu, ref, ref["iupp"]), parity.worst(gth, ref, ref["iupp"])
    record("co_kernel1", models=384, well_posed=int(wp.sum()), max_err_lu=float(w_lu[wp].max()),
           max_err_gth=float(w_gth[wp].max()), max_lu_vs_gth=float(parity.worst(lu, gth, ref["iupp"])[wp].max()))
    assert w_lu[wp].max() < RTOL and w_gth[wp].max() < RTOL
    assert parity.worst(lu, gth, ref["iupp"])[wp].max() < RTOL
    assert ((lu["status"] ^ gth["status"])[wp] & 8 == 0).all()
    # the models whose calls keep returning every level on the floor (NaN escape probabilities: their populations do not sum
    # to 1 in the reference either): the default kernels skip the elimination of those calls, kernel = 1 runs its LU on the
    # NaN matrix like the reference -- same un-normalised populations
    off = (np.abs(ref["xpop"].sum(axis=1) - 1.0) > 1e-6) & wp
    assert off.sum() >= 5, off.sum()
    for got in (lu, gth):
        np.testing.assert_allclose(got["xpop"][off].sum(axis=1), ref["xpop"][off].sum(axis=1), rtol=1e-6)
    np.testing.assert_allclose(gth["xpop"][off], lu["xpop"][off], rtol=1e-5, atol=1e-9 * 1e-5)
    # out-of-range inputs are refused the same way
    bad = gpu_solve(co_ctx, [0.0, 50.0], [1e4, 1e4], [1e15, 1e30], 2.7315, kernel=1)
    assert bad["status"][0] & 1 and bad["status"][1] & 2 and np.isnan(bad["surf"]).all()


def test_radex_class_with_other_colliders():
    """pyradex surface on the second table: collider keys are case-insensitive, H2 / e are what the file offers, an
    unknown or absent collider is a ValueError (core.py:492-512, base_class.py:224-263)."""
    import warnings
    from oracle.oracle import Oracle
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R = Radex(species=ROTOR, collider_densities={"h2": 3e4, "E": 5.0}, column=2e13, temperature=45.0,
                  tbackground=2.7315, deltav=1.0)
    assert [c.lower() for c in R.valid_colliders] == ["h2", "e"]
    R.run_radex()
    o = Oracle(ROTOR)
    ref = o.solve_batch_dens([45.0], [[3e4, 0, 0, 5.0, 0, 0, 0]], [2e13], tbg=2.7315)
    np.testing.assert_allclose(R.tex[:8], ref["tex"][0][:8], rtol=RTOL)
    np.testing.assert_allclose(R.tau[:8], ref["tau"][0][:8], rtol=RTOL)
    np.testing.assert_allclose(R.source_line_surfbrightness[:8], ref["surf"][0][:8], rtol=RTOL)
    # the drivers' {'oH2', 'pH2'} on a file whose partner is H2 itself: folded into H2 (core.py:551-556)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R.density = {"oH2": 7.5e3, "pH2": 2.5e3}
    R.run_radex()
    ref = o.solve_batch_dens([45.0], [[1e4, 0, 0, 0, 0, 0, 0]], [2e13], tbg=2.7315)
    np.testing.assert_allclose(R.tex[:8], ref["tex"][0][:8], rtol=RTOL)
    with pytest.raises(ValueError):
        R.density = {"He": 1e3}            # not a partner of this file: all-zero colliders
    with pytest.raises(ValueError):
        R.density = {"Xe": 1.0}


def test_lnprob_on_the_general_path(rotor_ctx, co_ctx):
    """k_lnprob_v1<1|2>: (a) the second molecule (total n goes to its H2 partner, as pyradex folds the drivers'
    oH2/pH2) against the oracle's lnprob; (b) CO through kernel=1 against the default fused kernel; (c) a file without
    any H2 partner is an argument error, not a silently wrong likelihood."""
    from oracle.oracle import Oracle
    L = _lib.load()
    rng = np.random.default_rng(12)
    o = Oracle(ROTOR)
    tbg = 2.7315 * 1.5
    jup = np.array([2, 3, 5, 7])
    truth = np.array([3.0, 1.7, 14.0, -10.2])
    n0 = 10 ** truth[0]
    surf = o.solve_batch_dens([10 ** truth[1]], [[n0, 0, 0, 0, 0, 0, 0]], [10 ** truth[2]], tbg=tbg)["surf"][0]
    flux = surf[jup - 1] * 10 ** truth[3] * 1e23 * (1 + 0.05 * rng.standard_normal(4))
    eflux = 0.1 * np.abs(flux)
    bounds1 = np.array([[2.0, 7.0], [np.log10(tbg), 3.0], [11.0, 16.5], [-14.0, -6.0]])
    P1 = np.vstack([truth + 0.2 * rng.standard_normal((150, 4)), rng.uniform(bounds1[:, 0] - 0.1, bounds1[:, 1] + 0.1, (50, 4))])
    obs = _lib.make_obs(jup, flux, eflux)
    out, ns = np.empty(200), C.c_int64(0)
    opts = _lib.default_opts()
    _lib.check(L.rb_lnprob1(rotor_ctx.handle, 200, _lib.ptr(P1), C.byref(obs), _lib.ptr(bounds1), tbg, C.byref(opts),
                            _lib.ptr(out), C.byref(ns)))
    ref = np.array([o.lnprob1(p, jup, flux, eflux, bounds1, tbg) for p in P1])
    assert ((out == -np.inf) == (ref == -np.inf)).all() and not np.isnan(out).any()
    fin = np.isfinite(ref)
    assert fin.sum() > 100 and ns.value == fin.sum() + ((ref == -np.inf) & np.isfinite(
        np.where((P1 > bounds1[:, 1]).any(axis=1) | (P1 < bounds1[:, 0]).any(axis=1) |
                 (P1[:, 2] - P1[:, 0] >= 17.5) | (P1[:, 2] - P1[:, 0] <= 10.0), -np.inf, 0.0))).sum()
    err = np.abs(out[fin] - ref[fin])
    assert (err < np.maximum(1e-4, 1e-7 * np.abs(ref[fin]))).all(), err.max()
    # two components
    bounds2 = np.vstack([bounds1, bounds1])
    bounds2[[0, 4], 0], bounds2[[2, 6], 0] = 1.5, 9.0
    t2 = np.array([3.2, 1.3, 13.0, -10.0, 3.9, 2.0, 14.0, -10.6])
    P2 = t2 + 0.15 * rng.standard_normal((120, 8))
    out2 = np.empty(120)
    _lib.check(L.rb_lnprob2(rotor_ctx.handle, 120, _lib.ptr(P2), C.byref(obs), _lib.ptr(bounds2), 1, 30.0, tbg,
                            C.byref(opts), _lib.ptr(out2), C.byref(ns)))
    ref2 = np.array([o.lnprob2(p, jup, flux, eflux, bounds2, 30.0, tbg) for p in P2])
    assert ((out2 == -np.inf) == (ref2 == -np.inf)).all()
    fin = np.isfinite(ref2)
    assert fin.sum() > 40
    err = np.abs(out2[fin] - ref2[fin])
    assert (err < np.maximum(1e-4, 1e-7 * np.abs(ref2[fin]))).all(), err.max()
    # (b) CO, kernel=1 against the default kernels
    from radex_emcee_b200 import emcee_radex as er1
    from radex_emcee_b200.data import get_source, read_data
    z, lw, j1, f1, e1 = get_source("G09v1.97", read_data(ROOT + "/data/flux.dat"))
    tb, ra, b1, p0 = er1.source_setup(z)
    Pc = p0 + 0.05 * rng.standard_normal((96, 4))
    obs1 = _lib.make_obs(j1, f1, e1)
    a, b = np.empty(96), np.empty(96)
    _lib.check(L.rb_lnprob1(co_ctx.handle, 96, _lib.ptr(Pc), C.byref(obs1), _lib.ptr(b1), tb, C.byref(opts), _lib.ptr(a), None))
    o1 = _lib.default_opts(kernel=1)
    _lib.check(L.rb_lnprob1(co_ctx.handle, 96, _lib.ptr(Pc), C.byref(obs1), _lib.ptr(b1), tb, C.byref(o1), _lib.ptr(b), None))
    fin = np.isfinite(a)
    assert fin.sum() > 80 and (np.isfinite(b) == fin).all()
    assert (np.abs(a[fin] - b[fin]) < np.maximum(1e-4, 1e-7 * np.abs(a[fin]))).all()
    # (c) no H2 / p-H2 / o-H2 partner at all
    txt = open(ROTOR).read()
    i0, i1 = txt.index("!COLLISIONS BETWEEN"), txt.rindex("!COLLISIONS BETWEEN")
    only_e = txt[:i0].replace("!NUMBER OF COLL PARTNERS\n2", "!NUMBER OF COLL PARTNERS\n1") + txt[i1:]
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "only_e.dat")
        open(path, "w").write(only_e)
        ctx_e = _lib.Context(_lib.MolData(path), 0)
        with pytest.raises(_lib.RadexB200Error, match="partner"):
            _lib.check(L.rb_lnprob1(ctx_e.handle, 8, _lib.ptr(P1[:8].copy()), C.byref(obs), _lib.ptr(bounds1), tbg,
                                    C.byref(opts), _lib.ptr(out[:8].copy()), None))
