"""Device stretch move: kernels vs their numpy restatement, same-seed chains vs the CPU reference sampler
(oracle lnprob), the in-library loop vs the per-half-step calls, sub-ensembles per source, NCCL ranks.
north_star: same-seed posterior medians within 0.01 dex."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import MOLFILE, ROOT
from radex_emcee_b200 import _lib
from radex_emcee_b200 import emcee_radex as er1
from radex_emcee_b200 import emcee_radex_2comp as er2
from radex_emcee_b200.data import get_source, read_data
from radex_emcee_b200.sampler import CudaEngine, SLEDModel, SplitSpec, StretchSampler
import ref_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _lib.Context(_lib.MolData(MOLFILE), 0)


def model1(name="G09v1.97"):
    z, lw, jup, flux, eflux = get_source(name, read_data(ROOT + "/data/flux.dat"))
    tbg, ra, bounds, p0 = er1.source_setup(z)
    return SLEDModel(1, jup, flux, eflux, bounds, tbg), p0


def model2(name="G09v1.97"):
    z, T_d, lw, jup, flux, eflux = get_source(name, read_data(ROOT + "/data/flux_for2p.dat"))
    tbg, ra, bounds, p0 = er2.source_setup(z)
    p0 = p0.copy()
    p0[3] += 0.1      # cold size > warm size so the whole starting ball has a finite prior
    return SLEDModel(2, jup, flux, eflux, bounds, tbg, T_d=T_d), p0


@pytest.fixture(scope="module")
def setup(ctx):
    m, p0 = model1()
    return ctx, m, p0


def test_kernels_match_numpy(setup):
    ctx, model, p0 = setup
    eng = CudaEngine(ctx, model)
    rng = np.random.default_rng(0)
    S = rng.standard_normal((1000, 4))
    Cc = rng.standard_normal((777, 4))
    Sd, Cd = torch.from_numpy(S).cuda(), torch.from_numpy(Cc).cuda()
    for step, half, gid0 in ((0, 0, 0), (5, 1, 1), (2 ** 33, 1, 2 ** 35 + 1)):
        Q, lf = eng.propose(Sd, Cd, 2.0, 1234567890123, step, half, gid0, 2)
        Qn, lfn, j, zz = ref_engine.propose_np(S, Cc, 2.0, 1234567890123, step, half, gid0, 2)
        np.testing.assert_array_equal(Q.cpu().numpy(), Qn)
        np.testing.assert_allclose(lf.cpu().numpy(), lfn, rtol=0, atol=1e-14)
        assert zz.min() >= 0.5 and zz.max() <= 2.0 and 0.9 < zz.mean() < 1.3
        assert len(np.unique(j)) > 400
    lnp = rng.standard_normal(1000)
    lnp_new = lnp + rng.standard_normal(1000)
    lnp_new[::17] = -np.inf
    Q, lf = eng.propose(Sd, Cd, 2.0, 99, 3, 0, 0, 2)
    S2, l2 = Sd.clone(), torch.from_numpy(lnp).cuda()
    nacc = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.accept(S2, l2, Q, torch.from_numpy(lnp_new).cuda(), lf, 99, 3, 0, 0, 2, nacc)
    Sn, ln = S.copy(), lnp.copy()
    acc = ref_engine.accept_np(Sn, ln, Q.cpu().numpy(), lnp_new, lf.cpu().numpy(), 99, 3, 0, 0, 2)
    assert int(nacc.item()) == int(acc.sum()) and 100 < acc.sum() < 900
    np.testing.assert_array_equal(S2.cpu().numpy(), Sn)
    np.testing.assert_array_equal(l2.cpu().numpy(), ln)


@pytest.mark.parametrize("N,W,B,randomize,gid_base,nlocal", [
    (100, 100, 100, True, 0, 100), (96, 48, 12, True, 48, 48), (4096, 4096, 512, True, 1024, 2048),
    (64, 64, 2, False, 32, 32), (1 << 16, 1 << 15, 1 << 12, True, 0, 1 << 16)])
def test_second_form_kernels_match_numpy(setup, N, W, B, randomize, gid_base, nlocal):
    """pack / propose2 / accept2 (split as a function of (seed, step, global id), partners from the walker's own
    sub-ensemble, per-walker acceptance counters, NaN count) against tests/ref_engine.py, bit for bit."""
    ctx, model, p0 = setup
    eng = CudaEngine(ctx, model)
    rng = np.random.default_rng(N + B)
    ndim = 4
    X = rng.standard_normal((nlocal, ndim))
    Call = rng.standard_normal((N // 2, ndim))
    sp = SplitSpec(N, W, B, randomize, 0xC0FFEE1234567)
    for step, half in ((0, 0), (7, 1), (2 ** 34 + 3, 0)):
        Xd = torch.from_numpy(X).cuda()
        Cg = eng.pack(sp, step, half, gid_base, Xd)
        np.testing.assert_array_equal(Cg.cpu().numpy(), ref_engine.pack_np(sp, step, half, gid_base, X))
        Q, lf, src = eng.propose2(sp, step, half, gid_base, Xd, torch.from_numpy(Call).cuda(), 2.0)
        Qn, lfn, srcn, _ = ref_engine.propose2_np(sp, step, half, gid_base, X, Call, 2.0)
        np.testing.assert_array_equal(Q.cpu().numpy(), Qn)
        np.testing.assert_allclose(lf.cpu().numpy(), lfn, rtol=0, atol=1e-14)
        np.testing.assert_array_equal(src.cpu().numpy(), srcn)
        lnp = rng.standard_normal(nlocal)
        lnp_new = rng.standard_normal(nlocal // 2)
        lnp_new[::11] = -np.inf
        lnp_new[5::29] = np.nan
        X2, l2 = Xd.clone(), torch.from_numpy(lnp).cuda()
        nacc = torch.zeros(nlocal, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        eng.accept2(sp, step, half, gid_base, X2, l2, Q, torch.from_numpy(lnp_new).cuda(), lf, nacc, cnt)
        Xn, ln, nn = X.copy(), lnp.copy(), np.zeros(nlocal, np.int64)
        acc, nnan = ref_engine.accept2_np(sp, step, half, gid_base, Xn, ln, Qn, lnp_new, lfn, nn)
        np.testing.assert_array_equal(X2.cpu().numpy(), Xn)
        np.testing.assert_array_equal(l2.cpu().numpy(), ln)
        np.testing.assert_array_equal(nacc.cpu().numpy(), nn)
        assert int(cnt[0].item()) == nnan == np.isnan(lnp_new).sum() and acc.sum() > nlocal // 20
    with pytest.raises(_lib.RadexB200Error, match="split"):
        eng.pack(SplitSpec(N, W, B, randomize, 1), 0, 0, gid_base + 1, torch.from_numpy(X).cuda())


@pytest.mark.parametrize("randomize", [True, False])
def test_chain_matches_cpu_reference_sampler(setup, oracle, randomize):
    """Config 1 shortened (64 walkers x 12 steps): device chain vs the same move on the CPU with
    the oracle's lnprob.  Identical RNG streams -> chains agree walker by walker except where a
    <=1e-4 lnprob difference flips an accept; medians must agree to 0.01 dex.  The in-library loop
    (CUDA graph of one step) and the per-half-step calls give the same chain bit for bit."""
    ctx, model, p0 = setup
    nw, nsteps = 64, 12
    pos = p0 + 1e-3 * np.random.default_rng(20170914).standard_normal((nw, 4))
    gpu = StretchSampler(nw, 4, CudaEngine(ctx, model), seed=42, randomize_split=randomize)
    assert gpu.native
    gpu.run_mcmc(pos, nsteps)
    py = StretchSampler(nw, 4, CudaEngine(ctx, model), seed=42, randomize_split=randomize, native=False)
    py.run_mcmc(pos, nsteps)
    np.testing.assert_array_equal(gpu.get_chain(), py.get_chain())
    np.testing.assert_array_equal(gpu.get_log_prob(), py.get_log_prob())
    np.testing.assert_array_equal(gpu.acceptance_fraction, py.acceptance_fraction)
    # 64 walkers: the in-library loop proposes the second half-step speculatively (rb_opts.spec_half = 0: automatic), one
    # lnprob launch of 3 x 32 candidates per step -- same chain (above), more solves; with spec_half = -1 it is the
    # sequential move launch for launch
    n_py = int(py.engine.total_solves.item())
    n_spec = gpu.total_solves + int(gpu.engine.total_solves.item())
    assert n_py <= n_spec <= 1.5 * n_py + nw, (n_py, n_spec)
    seq_model = SLEDModel(1, model.Jup, model.flux, model.eflux, model.bounds, model.tbg, opts=_lib.default_opts(spec_half=-1))
    seq = StretchSampler(nw, 4, CudaEngine(ctx, seq_model), seed=42, randomize_split=randomize)
    assert seq.native
    seq.run_mcmc(pos, nsteps)
    np.testing.assert_array_equal(seq.get_chain(), py.get_chain())
    np.testing.assert_array_equal(seq.get_log_prob(), py.get_log_prob())
    assert seq.total_solves + int(seq.engine.total_solves.item()) == n_py
    cpu = StretchSampler(nw, 4, ref_engine.NumpyEngine(
        ref_engine.oracle_lnprob1(oracle, model.Jup, model.flux, model.eflux, model.bounds, model.tbg)), seed=42,
        randomize_split=randomize)
    cpu.run_mcmc(pos, nsteps)
    cg, cc = gpu.get_chain(), cpu.get_chain()
    assert cg.shape == (nsteps, nw, 4)
    same = np.isclose(cg, cc, rtol=0, atol=1e-9).all(axis=2)
    assert same.mean() > 0.95, same.mean()
    assert np.abs(np.median(cg[-1], axis=0) - np.median(cc[-1], axis=0)).max() < 0.01
    assert np.abs(gpu.acceptance_fraction - cpu.acceptance_fraction).mean() < 0.05
    lg, lc = gpu.get_log_prob(), cpu.get_log_prob()
    assert np.abs(lg - lc)[same].max() < 1e-4
    # burn-in idiom of the drivers (emcee_radex.py:490-494): reset keeps the state; thin stores every 2nd step
    gpu.reset()
    gpu.run_mcmc(None, 6, thin=2)
    assert gpu.get_chain().shape == (3, nw, 4)
    py.reset()
    py.run_mcmc(None, 6, thin=2)
    np.testing.assert_array_equal(gpu.get_chain(), py.get_chain())


def test_two_component_chain_with_dust_prior(ctx, oracle):
    """emcee_radex_2comp.py's ensemble (8 parameters, T_cold ~ N(T_d, T_d), T_cold < T_warm, size_cold >= size_warm):
    48 walkers x 8 steps vs the CPU reference sampler with the oracle's two-component lnprob."""
    model, p0 = model2()
    nw, nsteps = 48, 8
    pos = p0 + 1e-3 * np.random.default_rng(7).standard_normal((nw, 8))
    gpu = StretchSampler(nw, 8, CudaEngine(ctx, model), seed=5)
    gpu.run_mcmc(pos, nsteps)
    cpu = StretchSampler(nw, 8, ref_engine.NumpyEngine(
        ref_engine.oracle_lnprob2(oracle, model.Jup, model.flux, model.eflux, model.bounds, model.T_d, model.tbg)), seed=5)
    cpu.run_mcmc(pos, nsteps)
    cg, cc = gpu.get_chain(), cpu.get_chain()
    same = np.isclose(cg, cc, rtol=0, atol=1e-9).all(axis=2)
    assert same.mean() > 0.9, same.mean()
    assert np.abs(np.median(cg[-1], axis=0) - np.median(cc[-1], axis=0)).max() < 0.01
    assert np.abs(gpu.get_log_prob() - cpu.get_log_prob())[same].max() < 1e-4
    assert 0.2 < gpu.acceptance_fraction.mean() < 0.95


def test_sources_fitted_concurrently(ctx):
    """BASELINE.json configs[3]: several flux.dat sources in ONE ensemble (own background temperature, bounds and line
    set per walker): every sub-ensemble's chain is bit for bit the chain of that source sampled alone."""
    data = read_data(ROOT + "/data/flux.dat")
    names = [n for n in data][:5]
    W, nsteps = 40, 6
    models, starts = [], []
    for k, nm in enumerate(names):
        m, p0 = model1(nm)
        models.append(m)
        starts.append(p0 + 1e-3 * np.random.default_rng(100 + k).standard_normal((W, 4)))
    assert len({m.tbg for m in models}) == len(models) and len({len(m.Jup) for m in models}) > 1
    allsrc = StretchSampler(W * len(names), 4, CudaEngine(ctx, models), seed=9, nsources=len(names))
    allsrc.run_mcmc(np.vstack(starts), nsteps)
    chain, lnp = allsrc.get_chain(), allsrc.get_log_prob()
    assert np.isfinite(lnp).all()
    for k, m in enumerate(models):
        # alone, the sub-ensemble has global ids 0..W-1: give it the ids (and split blocks) it has in the big ensemble
        # by sampling it as source k of an ensemble whose other sources are copies of itself
        ref = StretchSampler(W * len(names), 4, CudaEngine(ctx, [m] * len(names)), seed=9, nsources=len(names), native=False)
        ref.run_mcmc(np.vstack([starts[k]] * len(names)), nsteps)
        np.testing.assert_array_equal(ref.get_chain()[:, k * W:(k + 1) * W], chain[:, k * W:(k + 1) * W])
        np.testing.assert_array_equal(ref.get_log_prob()[:, k * W:(k + 1) * W], lnp[:, k * W:(k + 1) * W])
    # and each source's lnprob is the single-source entry point's
    er1.R = None
    for k, m in enumerate(models):
        R = er1.init_radex(m.tbg)
        R.set_params(tbg=m.tbg)
        got = er1.lnprob(chain[-1, k * W:(k + 1) * W], m.Jup, m.flux, m.eflux, bounds=m.bounds)
        np.testing.assert_array_equal(got, lnp[-1, k * W:(k + 1) * W])


def test_full_length_config1_posterior(setup):
    """Config 1 at the reference's full length (100 walkers, 100 burn + 500 steps, 1e-3 ball around p0, emcee's
    default randomized split): device chain vs the CPU reference sampler (oracle lnprob, 60 000 solves on the host
    cores), same seed.  north_star: posterior medians within 0.01 dex."""
    ctx, model, p0 = setup
    nw = 100
    pos = p0 + 1e-3 * np.random.RandomState(20170914).randn(nw, 4)
    gpu = StretchSampler(nw, 4, CudaEngine(ctx, model), seed=20170914)
    gpu.run_mcmc(pos, 100, store=False)
    gpu.reset()
    gpu.run_mcmc(None, 500)
    fn = ref_engine.oracle_lnprob1_threads(MOLFILE, model.Jup, model.flux, model.eflux, model.bounds, model.tbg)
    cpu = StretchSampler(nw, 4, ref_engine.NumpyEngine(fn), seed=20170914)
    cpu.run_mcmc(pos, 100, store=False)
    cpu.reset()
    cpu.run_mcmc(None, 500)
    fg, fc = gpu.get_chain(flat=True), cpu.get_chain(flat=True)
    med_g, med_c = np.median(fg, axis=0), np.median(fc, axis=0)
    assert np.abs(med_g - med_c).max() < 0.01, (med_g, med_c)
    # 16 / 84 percentiles (the drivers' error bars) to 0.02 dex
    assert np.abs(np.percentile(fg, [16, 84], axis=0) - np.percentile(fc, [16, 84], axis=0)).max() < 0.02
    assert abs(gpu.acceptance_fraction.mean() - cpu.acceptance_fraction.mean()) < 0.01
    same = np.isclose(gpu.get_chain(), cpu.get_chain(), rtol=0, atol=1e-9).all(axis=2)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "config1_posterior.txt"), "w") as f:
            f.write("config 1, 100 walkers x (100 burn + 500) steps, seed 20170914\nmedians gpu %s\nmedians cpu %s\n"
                    "identical walker-steps %.4f\nacceptance gpu %.4f cpu %.4f\n"
                    % (med_g, med_c, same.mean(), gpu.acceptance_fraction.mean(), cpu.acceptance_fraction.mean()))


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RB_ROOT"], "tests"))
from test_gpu_sampler import model2
from radex_emcee_b200 import _lib
from radex_emcee_b200.sampler import CudaEngine, StretchSampler
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
model, p0 = model2()
nw = 4096
pos = p0 + 0.02 * np.random.default_rng(3).standard_normal((nw, 8))
ctx = _lib.Context(_lib.MolData(os.path.join(os.environ["RB_ROOT"], "radex_emcee_b200", "data", "co.dat")), lr)
s = StretchSampler(nw, 8, CudaEngine(ctx, model), seed=11)
s.run_mcmc(pos, 4)
chain, lnp, acc = s.get_chain(), s.get_log_prob(), s.acceptance_fraction
if dist.get_rank() == 0:
    np.savez(os.environ["RB_OUT"], chain=chain, lnp=lnp, acc=acc)
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_nccl_two_ranks_match_single_process(ctx, tmp_path):
    """The multi-GPU path north_star names: walkers sharded over ranks, NCCL all-gather of the complementary half per
    half-step.  torchrun with two ranks == the single-process chain, bit for bit."""
    model, p0 = model2()
    nw = 4096
    pos = p0 + 0.02 * np.random.default_rng(3).standard_normal((nw, 8))
    ref = StretchSampler(nw, 8, CudaEngine(ctx, model), seed=11)
    ref.run_mcmc(pos, 4)
    out = str(tmp_path / "nccl2.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RB_ROOT=ROOT, RB_OUT=out)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                    "127.0.0.1", "--master-port", "29517", str(script)], check=True, env=env, timeout=600)
    got = np.load(out)
    np.testing.assert_array_equal(got["chain"], ref.get_chain())
    np.testing.assert_array_equal(got["lnp"], ref.get_log_prob())
    np.testing.assert_array_equal(got["acc"], ref.acceptance_fraction)


def test_speculative_half_step_with_several_sources(ctx):
    """rb_opts.spec_half: the in-library loop proposes the second half-step speculatively (both candidates per walker, one
    lnprob launch of 1.5 N walkers per step).  Four sources x 24 walkers, randomized split: the chain, the log-probabilities
    and the per-walker acceptance counters equal those of the sequential move (spec_half = -1) bit for bit."""
    data = read_data(ROOT + "/data/flux.dat")
    chains = []
    for spec in (0, -1, 1):
        models, starts = [], []
        for k, nm in enumerate(list(data)[:4]):
            z, lw, jup, flux, eflux = get_source(nm, data)
            tbg, ra, bounds, p0 = er1.source_setup(z)
            models.append(SLEDModel(1, jup, flux, eflux, bounds, tbg, opts=_lib.default_opts(spec_half=spec)))
            starts.append(p0 + 0.02 * np.random.default_rng(11 + k).standard_normal((24, 4)))
        s = StretchSampler(96, 4, CudaEngine(ctx, models), seed=9, nsources=4)
        assert s.native
        s.run_mcmc(np.vstack(starts), 15)
        chains.append((s.get_chain(), s.get_log_prob(), s.acceptance_fraction))
    for c in chains[1:]:
        for a, b in zip(chains[0], c):
            np.testing.assert_array_equal(a, b)
