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


# ---- second form (csrc/stretch.cuh): split as a function of (seed, step, global walker id) ---------------
def split_perm(x, B, keys):
    """Keyed bijection of [0, B) (st2::split_perm): x uint64 array of values < B, keys (k0..k3) per element."""
    w = 1
    while (1 << w) < B:
        w += 1
    mask = np.uint64((1 << w) - 1)
    s1, s2 = np.uint64((w + 1) >> 1), np.uint64(w // 3 if w >= 3 else 1)
    mult = [np.uint64(m) for m in (0x9E3779B1, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F)]
    x = np.array(x, dtype=np.uint64)
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        y = x[todo]
        for r in range(4):
            y = ((y * mult[r]) & MASK) + keys[r][todo] & mask
            y ^= y >> (s1 if r % 2 == 0 else s2)
        x[todo] = y
        todo[todo] = y >= np.uint64(B)
    return x


def slot_walker(split, step, half, gid_base, k):
    """Local walker index of the local slots k (array) of half `half` (st2::slot_walker)."""
    B = split.block
    hb = B // 2
    k = np.asarray(k, dtype=np.int64)
    lb, t = k // hb, k % hb
    if not split.randomize:
        return lb * B + 2 * t + half
    gb = (gid_base // B + lb).astype(np.uint64)
    st = np.full(gb.shape, step, dtype=np.uint64)
    k0 = (split.seed & 0xFFFFFFFF) ^ 0x53504C54
    k1 = ((split.seed >> 32) & 0xFFFFFFFF) ^ 0x72616E64
    keys = philox4x32_10(gb & MASK, gb >> np.uint64(32), st & MASK, st >> np.uint64(32), k0, k1)
    return lb * B + split_perm((half * hb + t).astype(np.uint64), B, keys).astype(np.int64)


def pack_np(split, step, half, gid_base, X):
    return X[slot_walker(split, step, half, gid_base, np.arange(X.shape[0] // 2))]


def propose2_np(split, step, half, gid_base, X, Call, a):
    nh, ndim = X.shape[0] // 2, X.shape[1]
    i = slot_walker(split, step, half, gid_base, np.arange(nh))
    gid = (gid_base + i).astype(np.uint64)
    r = philox4x32_10(*_counters(gid, step, half, False), split.seed & 0xFFFFFFFF, split.seed >> 32)
    u = u01(r[0], r[1])
    sq = (a - 1.0) * u + 1.0
    z = sq * sq / a
    W = split.walkers_per_source
    nc = W // 2
    src = (gid // np.uint64(W)).astype(np.int64)
    j = np.minimum((u01(r[2], r[3]) * nc).astype(np.int64), nc - 1) + src * nc
    cj = Call[j]
    Q = cj - (cj - X[i]) * z[:, None]
    return Q, (ndim - 1.0) * np.log(z), src.astype(np.int32), i


def accept2_np(split, step, half, gid_base, X, lnp, Q, lnp_new, logfac, naccept):
    nh = X.shape[0] // 2
    i = slot_walker(split, step, half, gid_base, np.arange(nh))
    gid = (gid_base + i).astype(np.uint64)
    r = philox4x32_10(*_counters(gid, step, half, True), split.seed & 0xFFFFFFFF, split.seed >> 32)
    with np.errstate(divide="ignore", invalid="ignore"):
        lnu = np.log(u01(r[0], r[1]))
        acc = (logfac + lnp_new - lnp[i]) > lnu
    X[i[acc]] = Q[acc]
    lnp[i[acc]] = lnp_new[acc]
    naccept[i[acc]] += 1
    return acc, int(np.isnan(lnp_new).sum())


class NumpyEngine:
    """lnprob_fn: callable (n, ndim) ndarray -> (n,) ndarray; with several sources a list of them, one per source."""

    def __init__(self, lnprob_fn):
        self.device = torch.device("cpu")
        self.lnprob_fn = lnprob_fn
        self.launches = 0
        self.total_solves = torch.zeros(1, dtype=torch.int64)

    def lnprob(self, P, src_id=None):
        P = P.numpy()
        if isinstance(self.lnprob_fn, (list, tuple)):
            sid = src_id.numpy()
            out = np.empty(P.shape[0])
            for s, fn in enumerate(self.lnprob_fn):
                m = sid == s
                if m.any():
                    out[m] = fn(P[m])
            return torch.from_numpy(out)
        return torch.from_numpy(np.asarray(self.lnprob_fn(P), dtype=np.float64))

    def pack(self, split, step, half, gid_base, X):
        return torch.from_numpy(pack_np(split, step, half, gid_base, X.numpy()))

    def propose2(self, split, step, half, gid_base, X, Call, a):
        Q, lf, src, _ = propose2_np(split, step, half, gid_base, X.numpy(), Call.numpy(), a)
        return torch.from_numpy(Q), torch.from_numpy(lf), torch.from_numpy(src)

    def accept2(self, split, step, half, gid_base, X, lnp, Q, lnp_new, logfac, naccept, counters):
        _, nn = accept2_np(split, step, half, gid_base, X.numpy(), lnp.numpy(), Q.numpy(), lnp_new.numpy(), logfac.numpy(),
                           naccept.numpy())
        counters[0] += nn

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


def oracle_lnprob1_threads(molfile, jup, flux, eflux, bounds, tbg, nthreads=None, **kw):
    """oracle_lnprob1 over the host cores: one Oracle (one RADEX COMMON-block state) per thread; ctypes releases the GIL."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle
    nthreads = nthreads or min(32, os.cpu_count() or 1)
    oracles = [Oracle(molfile) for _ in range(nthreads)]
    pool = ThreadPoolExecutor(nthreads)

    def fn(P):
        P = np.atleast_2d(P)
        chunks = np.array_split(np.arange(P.shape[0]), nthreads)

        def work(t):
            return [oracles[t].lnprob1(P[i], jup, flux, eflux, bounds, tbg, **kw) for i in chunks[t]]
        return np.array([v for part in pool.map(work, range(nthreads)) for v in part])
    return fn
