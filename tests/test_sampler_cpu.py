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


def run_chain(nw, nsteps, seed):
    rng = np.random.default_rng(5)
    p0 = rng.standard_normal((nw, 3))
    s = StretchSampler(nw, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=seed)
    s.run_mcmc(p0, nsteps)
    return s


def test_single_process_statistics():
    s = run_chain(64, 400, 11)
    c = s.get_chain()[100:].reshape(-1, 3)
    assert np.abs(c.mean(axis=0)).max() < 0.25
    np.testing.assert_allclose(c.std(axis=0), [1.0, 2.0, 0.5], rtol=0.2)
    assert 0.3 < s.acceptance_fraction < 0.9
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


def _worker(rank, world, port, nw, nsteps, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        p0 = rng.standard_normal((nw, 3))
        s = StretchSampler(nw, 3, ref_engine.NumpyEngine(gauss_lnprob), seed=seed)
        assert s.world == world and s.nlocal == nw // world
        s.run_mcmc(p0, nsteps)
        chain, lnp, acc = s.get_chain(), s.get_log_prob(), s.acceptance_fraction
        last_x, last_l = s.get_last_sample()
        if rank == 0:
            np.savez(out, chain=chain, lnp=lnp, acc=acc, last_x=last_x, last_l=last_l)
    finally:
        dist.destroy_process_group()


def test_world_size_2_matches_single_process(tmp_path):
    nw, nsteps, seed = 32, 25, 77
    ref = run_chain(nw, nsteps, seed)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, port, nw, nsteps, seed, out), nprocs=2, join=True)
    got = np.load(out)
    np.testing.assert_array_equal(got["chain"], ref.get_chain())
    np.testing.assert_array_equal(got["lnp"], ref.get_log_prob())
    assert got["acc"] == ref.acceptance_fraction
    x, l = ref.get_last_sample()
    np.testing.assert_array_equal(got["last_x"], x)
    np.testing.assert_array_equal(got["last_l"], l)
