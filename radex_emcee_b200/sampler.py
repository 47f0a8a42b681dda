"""Device-resident affine-invariant ensemble sampler (stretch move).

Replaces ``emcee.EnsembleSampler`` + its default ``StretchMove(a=2)`` for the drivers' call
sites (emcee/emcee_radex.py:483-499, emcee/emcee_radex_2comp.py:557-574): walkers,
log-probabilities, proposals and the accept test live on the GPU; ``lnprob`` is the fused
kernel of libradex_b200; nothing crosses PCIe per step unless the chain is stored on the host.

Red/blue split and sharding
  The ensemble is split by global walker id parity (even = half 0, odd = half 1), emcee's
  ``randomize_split=False`` variant of the same move.  With G ranks, rank r owns the contiguous
  global ids [r*N/G, (r+1)*N/G).  Per half-step each rank needs the *positions* of the whole
  complementary half: one ``all_gather_into_tensor`` (NCCL over NVLink on GPUs, gloo on CPU) of
  (N/2G) x ndim doubles per rank.  The Philox stream is keyed by (seed, step, half, global id),
  and gathered halves are ordered by global id, so the chain is independent of G.

The arithmetic is behind an ``engine`` object so that the host logic (sharding, gather order,
bookkeeping) can be exercised on CPU by the test-suite's engine (tests/ref_engine.py); the
package itself only ships the CUDA engine -- there is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class SLEDModel:
    """Everything the fused lnprob kernel needs for one source."""

    def __init__(self, ncomp, Jup, flux, eflux, bounds, tbg, T_d=None, opts=None):
        if ncomp not in (1, 2):
            raise ValueError("ncomp must be 1 or 2")
        self.ncomp = ncomp
        self.ndim = 4 * ncomp
        self.Jup = np.asarray(Jup, dtype=np.int64)
        self.flux = np.asarray(flux, dtype=np.float64)
        self.eflux = np.asarray(eflux, dtype=np.float64)
        self.bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        if self.bounds.shape != (self.ndim, 2):
            raise ValueError("bounds must have shape (%d, 2)" % self.ndim)
        self.tbg = float(tbg)
        self.T_d = None if T_d is None else float(T_d)
        self.opts = opts


class CudaEngine:
    """Stretch-move + lnprob arithmetic on one GPU through the C ABI (device pointers)."""

    def __init__(self, ctx: _lib.Context, model: SLEDModel):
        if not torch.cuda.is_available():
            raise _lib.RadexB200Error("CudaEngine needs a CUDA device; there is no CPU fallback")
        self.ctx = ctx
        self.model = model
        self.device = torch.device("cuda", ctx.device)
        self.L = _lib.load()
        self.obs = _lib.make_obs(model.Jup, model.flux, model.eflux)
        self.opts = model.opts if model.opts is not None else _lib.default_opts()
        self.nsolves = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.total_solves = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.launches = 0

    def _bind_stream(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def lnprob(self, P: torch.Tensor) -> torch.Tensor:
        self._bind_stream()
        n = P.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        m = self.model
        if m.ncomp == 1:
            rc = self.L.rb_lnprob1_dev(self.ctx.handle, n, P.data_ptr(), C.byref(self.obs), _lib.ptr(m.bounds), m.tbg,
                                       C.byref(self.opts), out.data_ptr(), self.nsolves.data_ptr())
        else:
            rc = self.L.rb_lnprob2_dev(self.ctx.handle, n, P.data_ptr(), C.byref(self.obs), _lib.ptr(m.bounds),
                                       int(m.T_d is not None), m.T_d if m.T_d is not None else 0.0, m.tbg,
                                       C.byref(self.opts), out.data_ptr(), self.nsolves.data_ptr())
        _lib.check(rc)
        self.total_solves += self.nsolves
        self.launches += 1
        return out

    def propose(self, S, Cpos, a, seed, step, half, gid0, gid_stride):
        self._bind_stream()
        ns, ndim = S.shape
        Q = torch.empty_like(S)
        logfac = torch.empty(ns, dtype=torch.float64, device=self.device)
        _lib.check(self.L.rb_stretch_propose_dev(self.ctx.handle, ns, ndim, S.data_ptr(), Cpos.shape[0], Cpos.data_ptr(),
                                                 a, seed, step, half, gid0, gid_stride, Q.data_ptr(), logfac.data_ptr()))
        self.launches += 1
        return Q, logfac

    def accept(self, S, lnp, Q, lnp_new, logfac, seed, step, half, gid0, gid_stride, naccept):
        self._bind_stream()
        ns, ndim = S.shape
        _lib.check(self.L.rb_stretch_accept_dev(self.ctx.handle, ns, ndim, S.data_ptr(), lnp.data_ptr(), Q.data_ptr(),
                                                lnp_new.data_ptr(), logfac.data_ptr(), seed, step, half, gid0,
                                                gid_stride, naccept.data_ptr()))
        self.launches += 1


class StretchSampler:
    """``EnsembleSampler``-like driver: ``run_mcmc``, ``get_chain``, ``get_log_prob``, ``reset``."""

    def __init__(self, nwalkers, ndim, engine, a=2.0, seed=0, group=None):
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        if nwalkers % (2 * self.world) != 0:
            raise ValueError("nwalkers must be a multiple of 2*world_size")
        if nwalkers < 2 * ndim:
            raise ValueError("emcee requires nwalkers >= 2*ndim")
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.engine = engine
        self.device = engine.device
        self.a = float(a)
        self.seed = int(seed)
        self.nlocal = self.nwalkers // self.world          # walkers owned by this rank
        self.nhalf = self.nlocal // 2                       # per half on this rank
        self.gid_base = self.rank * self.nlocal             # first global id owned (even)
        self.step = 0
        self.X = None                                       # [2][nhalf, ndim]
        self.lnp = None                                     # [2][nhalf]
        self.naccept = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.reset()

    # ---- storage --------------------------------------------------------------------------------
    def reset(self):
        self._chain, self._lnp_chain = [], []
        self.naccept.zero_()
        self.nsteps_done = 0

    def _gather_complement(self, Xc: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return Xc
        out = torch.empty((self.world * Xc.shape[0], Xc.shape[1]), dtype=Xc.dtype, device=Xc.device)
        self.dist.all_gather_into_tensor(out, Xc.contiguous(), group=self.group)
        return out

    def _local_from_global(self, p0: np.ndarray):
        loc = np.asarray(p0, dtype=np.float64)[self.gid_base:self.gid_base + self.nlocal]
        return [torch.from_numpy(np.ascontiguousarray(loc[h::2])).to(self.device) for h in (0, 1)]

    def set_state(self, p0):
        """p0: global (nwalkers, ndim) array, identical on every rank."""
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (self.nwalkers, self.ndim):
            raise ValueError("p0 must have shape (nwalkers, ndim)")
        self.X = self._local_from_global(p0)
        self.lnp = [self.engine.lnprob(x) for x in self.X]
        for l in self.lnp:
            if bool(torch.isnan(l).any()):      # emcee: "Probability function returned NaN"; -inf is allowed
                raise ValueError("Probability function returned NaN")

    # ---- the move -------------------------------------------------------------------------------
    def _half_step(self, half):
        S, lnp = self.X[half], self.lnp[half]
        Cfull = self._gather_complement(self.X[1 - half])
        gid0 = self.gid_base + half
        Q, logfac = self.engine.propose(S, Cfull, self.a, self.seed, self.step, half, gid0, 2)
        lnp_new = self.engine.lnprob(Q)
        self.engine.accept(S, lnp, Q, lnp_new, logfac, self.seed, self.step, half, gid0, 2, self.naccept)

    def run_mcmc(self, p0, nsteps, store=True, thin=1):
        if p0 is not None:
            self.set_state(p0)
        if self.X is None:
            raise ValueError("no initial state")
        for _ in range(int(nsteps)):
            self._half_step(0)
            self._half_step(1)
            self.step += 1
            self.nsteps_done += 1
            if store and (self.nsteps_done % thin == 0):
                self._chain.append(self._interleave(self.X).clone())
                self._lnp_chain.append(self._interleave(self.lnp).clone())
        return None

    def _interleave(self, halves):
        a, b = halves
        out = torch.empty((self.nlocal,) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
        out[0::2] = a
        out[1::2] = b
        return out

    # ---- results --------------------------------------------------------------------------------
    def _gather_all(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t
        shape = (self.world * t.shape[0],) + tuple(t.shape[1:])
        out = torch.empty(shape, dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def get_last_sample(self):
        """(positions (nwalkers, ndim), lnprob (nwalkers,)) gathered over ranks, as numpy."""
        x = self._gather_all(self._interleave(self.X))
        l = self._gather_all(self._interleave(self.lnp))
        return x.cpu().numpy(), l.cpu().numpy()

    def get_chain(self, flat=False):
        """(steps, nwalkers, ndim) like emcee v3's ``get_chain``."""
        if not self._chain:
            return np.empty((0, self.nwalkers, self.ndim))
        loc = torch.stack(self._chain)                       # steps, nlocal, ndim
        if self.world > 1:
            loc = self._gather_all(loc.transpose(0, 1).contiguous()).transpose(0, 1)
        c = loc.cpu().numpy()
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False):
        if not self._lnp_chain:
            return np.empty((0, self.nwalkers))
        loc = torch.stack(self._lnp_chain)
        if self.world > 1:
            loc = self._gather_all(loc.transpose(0, 1).contiguous()).transpose(0, 1)
        c = loc.cpu().numpy()
        return c.reshape(-1) if flat else c

    @property
    def acceptance_fraction(self):
        n = self.naccept.clone()
        if self.world > 1:
            self.dist.all_reduce(n, group=self.group)
        return float(n.item()) / max(1, self.nwalkers * self.nsteps_done)
