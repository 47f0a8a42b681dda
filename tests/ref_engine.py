"""Test-only CPU engine for StretchSampler: numpy restatement of the device stretch-move kernels
(same Philox4x32-10 stream, same roundings) with the CPU oracle as lnprob.  Lets the host logic
of the sampler (sharding, gather order) run under gloo on CPU, and is the reference the GPU
sampler is compared against.  Never imported by the package."""
import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0) & MASK
    k1 = np.uint64(k1) & MASK
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def u01(hi, lo):
    v = (hi << np.uint64(32)) | lo
    return (v >> np.uint64(11)).astype(np.float64) * 1.1102230246251565e-16


def _counters(gid, step, half, accept):
    sh = np.uint64(step * 2 + half)
    c2 = np.full(gid.shape, sh & MASK, dtype=np.uint64)
    c3v = (int(sh) >> 32) & 0x7FFFFFFF
    if accept:
        c3v |= 0x80000000
    c3 = np.full(gid.shape, c3v, dtype=np.uint64)
    return gid & MASK, gid >> np.uint64(32), c2, c3


def propose_np(S, C, a, seed, step, half, gid0, gid_stride):
    ns, ndim = S.shape
    gid = (np.uint64(gid0) + np.arange(ns, dtype=np.uint64) * np.uint64(gid_stride))
    r = philox4x32_10(*_counters(gid, step, half, False), seed & 0xFFFFFFFF, seed >> 32)
    u = u01(r[0], r[1])
    sq = (a - 1.0) * u + 1.0
    z = sq * sq / a
    j = np.minimum((u01(r[2], r[3]) * C.shape[0]).astype(np.int64), C.shape[0] - 1)
    cj = C[j]
    Q = cj - (cj - S) * z[:, None]
    return Q, (ndim - 1.0) * np.log(z), j, z


def accept_np(S, lnp, Q, lnp_new, logfac, seed, step, half, gid0, gid_stride):
    ns = S.shape[0]
    gid = (np.uint64(gid0) + np.arange(ns, dtype=np.uint64) * np.uint64(gid_stride))
    r = philox4x32_10(*_counters(gid, step, half, True), seed & 0xFFFFFFFF, seed >> 32)
    with np.errstate(divide="ignore", invalid="ignore"):
        lnu = np.log(u01(r[0], r[1]))
        acc = (logfac + lnp_new - lnp) > lnu
    S[acc] = Q[acc]
    lnp[acc] = lnp_new[acc]
    return acc


class NumpyEngine:
    """lnprob_fn: callable (n, ndim) ndarray -> (n,) ndarray."""

    def __init__(self, lnprob_fn):
        self.device = torch.device("cpu")
        self.lnprob_fn = lnprob_fn
        self.launches = 0

    def lnprob(self, P):
        return torch.from_numpy(np.asarray(self.lnprob_fn(P.numpy()), dtype=np.float64))

    def propose(self, S, Cpos, a, seed, step, half, gid0, gid_stride):
        Q, lf, _, _ = propose_np(S.numpy(), Cpos.numpy(), a, seed, step, half, gid0, gid_stride)
        return torch.from_numpy(Q), torch.from_numpy(lf)

    def accept(self, S, lnp, Q, lnp_new, logfac, seed, step, half, gid0, gid_stride, naccept):
        acc = accept_np(S.numpy(), lnp.numpy(), Q.numpy(), lnp_new.numpy(), logfac.numpy(), seed, step, half, gid0,
                        gid_stride)
        naccept += int(acc.sum())


def oracle_lnprob1(oracle, jup, flux, eflux, bounds, tbg, **kw):
    def fn(P):
        return np.array([oracle.lnprob1(p, jup, flux, eflux, bounds, tbg, **kw) for p in np.atleast_2d(P)])
    return fn


def oracle_lnprob2(oracle, jup, flux, eflux, bounds, t_d, tbg, **kw):
    def fn(P):
        return np.array([oracle.lnprob2(p, jup, flux, eflux, bounds, t_d, tbg, **kw) for p in np.atleast_2d(P)])
    return fn
