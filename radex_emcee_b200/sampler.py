"""Device-resident affine-invariant ensemble sampler (stretch move).

Replaces ``emcee.EnsembleSampler`` + its default ``StretchMove(a=2)`` for the drivers' call
sites (emcee/emcee_radex.py:483-499, emcee/emcee_radex_2comp.py:557-574): walkers,
log-probabilities, proposals and the accept test live on the GPU; ``lnprob`` is the fused
kernel (or the solve pipeline) of libradex_b200; nothing crosses PCIe per step.

Red/blue split
  emcee's default ``randomize_split=True`` draws a new balanced random labelling of the walkers every
  step.  Here the labelling is a keyed bijection evaluated per BLOCK of ``split_block`` consecutive
  walkers from (seed, step, block) -- see ``csrc/stretch.cuh`` -- so every rank can evaluate it for any
  walker without communication.  ``randomize_split=False`` is the parity split (even ids / odd ids).

Sub-ensembles (BASELINE.json configs[3], all sources of flux.dat fitted concurrently)
  ``nsources`` independent ensembles of ``nwalkers // nsources`` walkers share one state array; a
  walker's partner is drawn from the complementary half of its own sub-ensemble and its log-probability
  is evaluated against its own source (``CudaEngine`` over several ``SLEDModel``s: one fused launch
  with a per-walker source row).

Sharding
  With G ranks, rank r owns the contiguous global ids [r N/G, (r+1) N/G).  Per half-step each rank
  needs the *positions* of the whole complementary half: one ``all_gather_into_tensor`` (NCCL over
  NVLink on GPUs, gloo on CPU) of (N/2G) x ndim doubles per rank.  The Philox streams are keyed by
  (seed, step, half, global id) and the gathered half is in global slot order, so the chain is
  independent of G.  On one GPU the whole loop runs inside the library (``rb_stretch_run_dev``; small
  ensembles replay a CUDA graph of one step) -- same kernels, same chain.

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
    """Everything the lnprob kernels need for one source."""

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


class SplitSpec:
    """The red/blue split of an ensemble (rb_split of the C ABI)."""

    def __init__(self, nwalkers, walkers_per_source, block, randomize, seed):
        self.nwalkers, self.walkers_per_source, self.block = int(nwalkers), int(walkers_per_source), int(block)
        self.randomize, self.seed = bool(randomize), int(seed)
        if self.nwalkers % self.walkers_per_source or self.walkers_per_source % self.block or self.block % 2:
            raise ValueError("need split_block | walkers per source | nwalkers and an even split_block")

    def c_struct(self):
        return _lib.rb_split(self.nwalkers, self.walkers_per_source, self.block, int(self.randomize), 0, self.seed)


def default_split_block(walkers_per_source, randomize):
    """Blocks the split is balanced over.  The chain depends on it, so the rule only looks at the
    ensemble: an eighth of a large sub-ensemble (any rank count up to 8 owns whole blocks), the whole
    sub-ensemble otherwise (exactly emcee's balanced random labelling)."""
    if not randomize:
        return 2
    W = int(walkers_per_source)
    return W // 8 if (W >= 1024 and W % 16 == 0) else W


class CudaEngine:
    """Stretch-move + lnprob arithmetic on one GPU through the C ABI (device pointers).
    ``model``: one SLEDModel, or a list of them (one per source, same ncomp)."""

    def __init__(self, ctx: _lib.Context, model):
        if not torch.cuda.is_available():
            raise _lib.RadexB200Error("CudaEngine needs a CUDA device; there is no CPU fallback")
        self.ctx = ctx
        self.models = list(model) if isinstance(model, (list, tuple)) else [model]
        self.model = self.models[0]
        if any(m.ncomp != self.model.ncomp for m in self.models):
            raise ValueError("all sources of one ensemble must have the same number of components")
        self.device = torch.device("cuda", ctx.device)
        self.L = _lib.load()
        self.obs = _lib.make_obs(self.model.Jup, self.model.flux, self.model.eflux)
        self.opts = self.model.opts if self.model.opts is not None else _lib.default_opts()
        self.srcset = _lib.SourceSet(ctx, self.model.ncomp,
                                     [_lib.make_source(m.ncomp, m.Jup, m.flux, m.eflux, m.bounds, m.tbg, m.T_d)
                                      for m in self.models])
        self.nsolves = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.total_solves = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.launches = 0

    def _bind_stream(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def lnprob(self, P: torch.Tensor, src_id=None) -> torch.Tensor:
        self._bind_stream()
        n = P.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        if len(self.models) > 1 and src_id is None:
            raise ValueError("src_id is required with more than one source")
        _lib.check(self.L.rb_lnprob_src_dev(self.ctx.handle, self.srcset.handle, n, P.data_ptr(),
                                            src_id.data_ptr() if src_id is not None else None, C.byref(self.opts),
                                            out.data_ptr(), self.nsolves.data_ptr()))
        self.total_solves += self.nsolves
        self.launches += 1
        return out

    # ---- first form of the C ABI (separate half arrays, parity split) -------------------------------
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

    # ---- second form: the ensemble stays in place -----------------------------------------------------
    def pack(self, split, step, half, gid_base, X):
        self._bind_stream()
        nlocal, ndim = X.shape
        out = torch.empty((nlocal // 2, ndim), dtype=torch.float64, device=self.device)
        sp = split.c_struct()
        _lib.check(self.L.rb_stretch_pack_dev(self.ctx.handle, C.byref(sp), step, half, gid_base, nlocal, ndim,
                                              X.data_ptr(), out.data_ptr()))
        self.launches += 1
        return out

    def propose2(self, split, step, half, gid_base, X, Call, a):
        self._bind_stream()
        nlocal, ndim = X.shape
        Q = torch.empty((nlocal // 2, ndim), dtype=torch.float64, device=self.device)
        logfac = torch.empty(nlocal // 2, dtype=torch.float64, device=self.device)
        src_id = torch.empty(nlocal // 2, dtype=torch.int32, device=self.device)
        sp = split.c_struct()
        _lib.check(self.L.rb_stretch_propose2_dev(self.ctx.handle, C.byref(sp), step, half, gid_base, nlocal, ndim,
                                                  X.data_ptr(), Call.data_ptr(), a, Q.data_ptr(), logfac.data_ptr(),
                                                  src_id.data_ptr()))
        self.launches += 1
        return Q, logfac, src_id

    def accept2(self, split, step, half, gid_base, X, lnp, Q, lnp_new, logfac, naccept, counters):
        self._bind_stream()
        nlocal, ndim = X.shape
        sp = split.c_struct()
        _lib.check(self.L.rb_stretch_accept2_dev(self.ctx.handle, C.byref(sp), step, half, gid_base, nlocal, ndim,
                                                 X.data_ptr(), lnp.data_ptr(), Q.data_ptr(), lnp_new.data_ptr(),
                                                 logfac.data_ptr(), naccept.data_ptr(), counters.data_ptr()))
        self.launches += 1

    def run_native(self, split, a, step0, nsteps, X, lnp, naccept, counters, thin, chain, lnp_chain):
        """nsteps steps of the whole (single-rank) ensemble inside the library."""
        self._bind_stream()
        sp = split.c_struct()
        _, l0 = self.ctx.counters()
        _lib.check(self.L.rb_stretch_run_dev(self.ctx.handle, self.srcset.handle, C.byref(sp), a, step0, nsteps,
                                             C.byref(self.opts), X.data_ptr(), lnp.data_ptr(), naccept.data_ptr(),
                                             counters.data_ptr(), thin,
                                             chain.data_ptr() if chain is not None else None,
                                             lnp_chain.data_ptr() if lnp_chain is not None else None))
        self.launches += self.ctx.counters()[1] - l0


def walkers_independent(coords):
    """emcee's check of an initial ensemble (ensemble.py, ``walkers_independent``): finite, no degenerate
    dimension, condition number of the normalised, centred coordinates <= 1e8."""
    coords = np.asarray(coords, dtype=np.float64)
    if not np.all(np.isfinite(coords)):
        return False
    Cm = coords - np.mean(coords, axis=0)[None, :]
    colmax = np.amax(np.abs(Cm), axis=0)
    if np.any(colmax == 0):
        return False
    Cm = Cm / colmax
    Cm = Cm / np.sqrt(np.sum(Cm ** 2, axis=0))
    return bool(np.linalg.cond(Cm) <= 1e8)


class StretchSampler:
    """``EnsembleSampler``-like driver: ``run_mcmc``, ``get_chain``, ``get_log_prob``, ``reset``,
    ``acceptance_fraction`` (per walker, like emcee)."""

    def __init__(self, nwalkers, ndim, engine, a=2.0, seed=0, group=None, randomize_split=True, nsources=1,
                 split_block=None, native=True, time_gather=False):
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        nwalkers, nsources = int(nwalkers), int(nsources)
        if nwalkers % (2 * self.world) != 0:
            raise ValueError("nwalkers must be a multiple of 2*world_size")
        if nsources < 1 or nwalkers % nsources or (nwalkers // nsources) % 2:
            raise ValueError("nwalkers must be an even multiple of nsources")
        if nwalkers // nsources < 2 * ndim:
            raise ValueError("emcee requires nwalkers >= 2*ndim")
        self.nwalkers, self.ndim, self.nsources = nwalkers, int(ndim), nsources
        self.engine = engine
        self.device = engine.device
        self.a = float(a)
        self.seed = int(seed)
        self.nlocal = self.nwalkers // self.world          # walkers owned by this rank: global ids gid_base ...
        self.gid_base = self.rank * self.nlocal
        W = self.nwalkers // nsources
        block = int(split_block) if split_block else default_split_block(W, randomize_split)
        self.split = SplitSpec(self.nwalkers, W, block, randomize_split, self.seed)
        if self.nlocal % block:
            raise ValueError("every rank must own whole split blocks: nwalkers/world_size = %d is not a multiple of "
                             "split_block = %d (pass split_block, or randomize_split=False)" % (self.nlocal, block))
        self.native = bool(native) and self.world == 1 and hasattr(engine, "run_native")
        self.time_gather = bool(time_gather)
        self._gather_events = []
        self.step = 0
        self.X = None                                       # [nlocal, ndim]
        self.lnp = None                                     # [nlocal]
        self.naccept = torch.zeros(self.nlocal, dtype=torch.int64, device=self.device)
        self.counters = torch.zeros(2, dtype=torch.int64, device=self.device)     # NaN log-probabilities, solves
        self.reset()

    # ---- storage --------------------------------------------------------------------------------
    def reset(self):
        self._chain, self._lnp_chain = [], []
        self.naccept.zero_()
        self.nsteps_done = 0

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if self.time_gather and t.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            e1.record()
            self._gather_events.append((e0, e1))
        else:
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def gather_ms(self):
        """Device time spent in the all-gathers since the last call (needs time_gather=True)."""
        ms = sum(a.elapsed_time(b) for a, b in self._gather_events)
        self._gather_events = []
        return ms

    def _src_of_local(self):
        gid = self.gid_base + torch.arange(self.nlocal, device=self.device)
        return (gid // self.split.walkers_per_source).to(torch.int32)

    def set_state(self, p0, check=True):
        """p0: global (nwalkers, ndim) array, identical on every rank."""
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (self.nwalkers, self.ndim):
            raise ValueError("p0 must have shape (nwalkers, ndim)")
        W = self.split.walkers_per_source
        if check:
            for s in range(self.nsources):
                if not walkers_independent(p0[s * W:(s + 1) * W]):
                    raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                     "linearly independent for the best performance")
        loc = np.array(p0[self.gid_base:self.gid_base + self.nlocal], dtype=np.float64, order="C", copy=True)   # never alias the caller's array
        self.X = torch.from_numpy(loc).to(self.device)
        self.lnp = self.engine.lnprob(self.X, self._src_of_local() if self.nsources > 1 else None)
        if bool(torch.isnan(self.lnp).any()):      # emcee: "Probability function returned NaN"; -inf is allowed
            raise ValueError("Probability function returned NaN")

    # ---- the move -------------------------------------------------------------------------------
    def _half_step(self, half):
        eng, sp = self.engine, self.split
        Call = self._gather(eng.pack(sp, self.step, 1 - half, self.gid_base, self.X))
        Q, logfac, src_id = eng.propose2(sp, self.step, half, self.gid_base, self.X, Call, self.a)
        lnp_new = eng.lnprob(Q, src_id if self.nsources > 1 else None)
        eng.accept2(sp, self.step, half, self.gid_base, self.X, self.lnp, Q, lnp_new, logfac, self.naccept,
                    self.counters)

    def run_mcmc(self, p0, nsteps, store=True, thin=1):
        if p0 is not None:
            self.set_state(p0)
        if self.X is None:
            raise ValueError("no initial state")
        nsteps, thin = int(nsteps), int(thin)
        if self.native and nsteps > 0:
            # `thin` counts from this call's first step in the library; keep it aligned with nsteps_done
            nstore = (self.nsteps_done + nsteps) // thin - self.nsteps_done // thin if store else 0
            aligned = self.nsteps_done % thin == 0
            if store and nstore and aligned:
                chain = torch.empty((nstore, self.nlocal, self.ndim), dtype=torch.float64, device=self.device)
                lchain = torch.empty((nstore, self.nlocal), dtype=torch.float64, device=self.device)
            else:
                chain = lchain = None
            if not store or aligned:
                self.engine.run_native(self.split, self.a, self.step, nsteps, self.X, self.lnp, self.naccept,
                                       self.counters, thin, chain, lchain)
                self.step += nsteps
                self.nsteps_done += nsteps
                if chain is not None:
                    self._chain.append(chain)
                    self._lnp_chain.append(lchain)
                self._check_nan()
                return None
        for _ in range(nsteps):
            self._half_step(0)
            self._half_step(1)
            self.step += 1
            self.nsteps_done += 1
            if store and (self.nsteps_done % thin == 0):
                self._chain.append(self.X.clone()[None])
                self._lnp_chain.append(self.lnp.clone()[None])
        self._check_nan()
        return None

    def _check_nan(self):
        if int(self.counters[0].item()) > 0:
            raise ValueError("Probability function returned NaN")

    # ---- results --------------------------------------------------------------------------------
    def get_last_sample(self):
        """(positions (nwalkers, ndim), lnprob (nwalkers,)) gathered over ranks, as numpy."""
        return self._gather(self.X).cpu().numpy(), self._gather(self.lnp).cpu().numpy()

    def _stacked(self, parts, tail):
        if not parts:
            return np.empty((0, self.nwalkers) + tail)
        loc = torch.cat(parts)                               # steps, nlocal, ...
        if self.world > 1:
            loc = self._gather(loc.transpose(0, 1).contiguous()).transpose(0, 1)
        return loc.cpu().numpy()

    def get_chain(self, flat=False):
        """(steps, nwalkers, ndim) like emcee v3's ``get_chain``."""
        c = self._stacked(self._chain, (self.ndim,))
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False):
        c = self._stacked(self._lnp_chain, ())
        return c.reshape(-1) if flat else c

    @property
    def acceptance_fraction(self):
        """Per walker, like ``EnsembleSampler.acceptance_fraction`` (shape (nwalkers,))."""
        n = self._gather(self.naccept)
        return n.cpu().numpy().astype(np.float64) / max(1, self.nsteps_done)

    @property
    def total_solves(self):
        return int(self.counters[1].item())
