"""Host logic of the stretch-move sampler on CPU: Philox known answers, sharding / all-gather order
(world_size 2 over gloo reproduces the single-process chain bit for bit), basic statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ref_engine
from radex_emcee_b200.sampler import StretchSampler


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        r = ref_engine.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in r) == out


def gauss_lnprob(P):
    P = np.atleast_2d(P)
    return -0.5 * np.sum((P / np.array([1.0, 2.0, 0.5])) ** 2, axis=1)


def gauss_lnprob_shifted(P):
    P = np.atleast_2d(P)
    return -0.5 * np.sum(((P - np.array([5.0, -3.0, 1.0])) / np.array([0.3, 1.0, 2.0])) ** 2, axis=1)


def run_chain(nw, nsteps, seed, **kw):
    rng = np.random.default_rng(5)
    p0 = rng.standard_normal((nw, 3))
    s = StretchSampler(nw, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=seed, **kw)
    s.run_mcmc(p0, nsteps)
    return s


@pytest.mark.parametrize("randomize", [True, False])
def test_single_process_statistics(randomize):
    s = run_chain(64, 400, 11, randomize_split=randomize)
    c = s.get_chain()[100:].reshape(-1, 3)
    assert np.abs(c.mean(axis=0)).max() < 0.25
    np.testing.assert_allclose(c.std(axis=0), [1.0, 2.0, 0.5], rtol=0.2)
    af = s.acceptance_fraction                     # per walker, like emcee
    assert af.shape == (64,) and 0.3 < af.mean() < 0.9 and af.min() > 0.1
    assert s.get_log_prob().shape == (400, 64)
    # reset keeps the state, clears the stored chain (emcee's burn-in idiom, emcee_radex.py:490-494)
    s.reset()
    assert s.get_chain().shape[0] == 0
    s.run_mcmc(None, 3)
    assert s.get_chain().shape == (3, 64, 3)
    with pytest.raises(ValueError):
        StretchSampler(4, 3, ref_engine.NumpyEngine(gauss_lnprob))      # nwalkers < 2*ndim
    with pytest.raises(ValueError):
        StretchSampler(9, 3, ref_engine.NumpyEngine(gauss_lnprob))      # odd


def test_split_is_a_balanced_partition_that_changes_every_step():
    """emcee's randomize_split=True: every step a new labelling with nwalkers/2 walkers per half; here balanced per
    block and a function of (seed, step, block) only."""
    from radex_emcee_b200.sampler import SplitSpec, default_split_block
    assert default_split_block(100, True) == 100 and default_split_block(1 << 20, True) == 1 << 17
    assert default_split_block(100, False) == 2
    for N, W, B in ((100, 100, 100), (96, 48, 12), (4096, 4096, 512), (64, 64, 2)):
        sp = SplitSpec(N, W, B, True, 99)
        labels = []
        for step in range(6):
            i0 = ref_engine.slot_walker(sp, step, 0, 0, np.arange(N // 2))
            i1 = ref_engine.slot_walker(sp, step, 1, 0, np.arange(N // 2))
            assert sorted(np.concatenate([i0, i1]).tolist()) == list(range(N))          # a partition
            lab = np.zeros(N, int)
            lab[i1] = 1
            assert (lab.reshape(-1, B).sum(axis=1) == B // 2).all()                    # balanced per block
            # a rank that owns only the second half of the ensemble evaluates the same labelling
            if (N // 2) % B == 0:
                j0 = ref_engine.slot_walker(sp, step, 0, N // 2, np.arange(N // 4)) + N // 2
                np.testing.assert_array_equal(j0, i0[N // 4:])
            labels.append(lab)
        if B > 2:
            assert any((labels[0] != l).any() for l in labels[1:])
    sp = SplitSpec(64, 64, 2, False, 1)                                                 # parity split
    np.testing.assert_array_equal(ref_engine.slot_walker(sp, 7, 1, 0, np.arange(32)), 2 * np.arange(32) + 1)
    with pytest.raises(ValueError):
        SplitSpec(100, 100, 30, True, 0)


def test_sub_ensembles_do_not_mix():
    """Two sources in one state array (config 4): partners come from the walker's own sub-ensemble, each half of the
    state converges to its own target."""
    rng = np.random.default_rng(8)
    p0 = np.vstack([rng.standard_normal((32, 3)), rng.standard_normal((32, 3)) + np.array([5.0, -3.0, 1.0])])
    eng = ref_engine.NumpyEngine([gauss_lnprob, gauss_lnprob_shifted])
    s = StretchSampler(64, 3, eng, seed=3, nsources=2)
    assert s.split.walkers_per_source == 32 and s.split.block == 32
    s.run_mcmc(p0, 300)
    c = s.get_chain()[100:]
    np.testing.assert_allclose(c[:, :32].reshape(-1, 3).mean(axis=0), [0, 0, 0], atol=0.3)
    np.testing.assert_allclose(c[:, 32:].reshape(-1, 3).mean(axis=0), [5.0, -3.0, 1.0], atol=0.4)
    np.testing.assert_allclose(c[:, 32:].reshape(-1, 3).std(axis=0), [0.3, 1.0, 2.0], rtol=0.25)
    # the first sub-ensemble's chain is the chain it has when sampled alone (same seed, same global ids)
    alone = StretchSampler(32, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=3)
    alone.run_mcmc(p0[:32], 300)
    np.testing.assert_array_equal(alone.get_chain(), s.get_chain()[:, :32])


def test_nan_and_degenerate_start_are_errors():
    """emcee raises on a NaN log-probability and on a linearly dependent initial ensemble."""
    def bad(P):
        out = gauss_lnprob(P)
        out[np.atleast_2d(P)[:, 0] > 1.5] = np.nan
        return out
    rng = np.random.default_rng(1)
    p0 = 0.1 * rng.standard_normal((32, 3))
    s = StretchSampler(32, 3, ref_engine.NumpyEngine(bad), seed=1)
    with pytest.raises(ValueError, match="NaN"):
        s.run_mcmc(p0, 200)
    s = StretchSampler(32, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=1)
    p1 = p0.copy()
    p1[:, 2] = p1[:, 1]
    with pytest.raises(ValueError, match="condition number"):
        s.run_mcmc(p1, 1)
    p1[:, 2] = 0.0
    with pytest.raises(ValueError, match="condition number"):
        s.run_mcmc(p1, 1)


def _worker(rank, world, port, nw, nsteps, seed, out, kw):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        p0 = rng.standard_normal((nw, 3))
        s = StretchSampler(nw, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=seed, **kw)
        assert s.world == world and s.nlocal == nw // world
        s.run_mcmc(p0, nsteps)
        chain, lnp, acc = s.get_chain(), s.get_log_prob(), s.acceptance_fraction
        last_x, last_l = s.get_last_sample()
        if rank == 0:
            np.savez(out, chain=chain, lnp=lnp, acc=acc, last_x=last_x, last_l=last_l)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kw", [dict(randomize_split=True, split_block=8), dict(randomize_split=False)])
def test_world_size_2_matches_single_process(tmp_path, kw):
    nw, nsteps, seed = 32, 25, 77
    ref = run_chain(nw, nsteps, seed, **kw)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, port, nw, nsteps, seed, out, kw), nprocs=2, join=True)
    got = np.load(out)
    np.testing.assert_array_equal(got["chain"], ref.get_chain())
    np.testing.assert_array_equal(got["lnp"], ref.get_log_prob())
    np.testing.assert_array_equal(got["acc"], ref.acceptance_fraction)
    x, l = ref.get_last_sample()
    np.testing.assert_array_equal(got["last_x"], x)
    np.testing.assert_array_equal(got["last_l"], l)


def test_split_bijection_property():
    """For any block size, seed and step the keyed map is a bijection of [0, block) (hypothesis)."""
    from hypothesis import given, settings, strategies as st
    from radex_emcee_b200.sampler import SplitSpec

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 300).map(lambda k: 2 * k), st.integers(0, 2 ** 64 - 1), st.integers(0, 2 ** 40), st.integers(0, 5))
    def check(block, seed, step, blk):
        sp = SplitSpec(block * 6, block * 6, block, True, seed)
        i0 = ref_engine.slot_walker(sp, step, 0, blk * block, np.arange(block // 2))
        i1 = ref_engine.slot_walker(sp, step, 1, blk * block, np.arange(block // 2))
        assert sorted(np.concatenate([i0, i1]).tolist()) == list(range(block))

    check()
