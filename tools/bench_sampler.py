#!/usr/bin/env python
"""Walker-steps/s of the device-resident stretch-move sampler (BASELINE.json configs 1/3/5).

  python tools/bench_sampler.py --ncomp 2 --log2w 14 --steps 20          (1 GPU)
  torchrun --nproc-per-node N ... tools/bench_sampler.py --ncomp 2 --log2w 20 --steps 10

Walkers start in a small ball around p0 of the G09v1.97 setup (emcee_radex*.py main()); each step is two
half-steps = nwalkers fused lnprob evaluations.  Prints one JSON line on rank 0.
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from radex_emcee_b200 import _lib, emcee_radex as er1, emcee_radex_2comp as er2
from radex_emcee_b200.data import get_source, read_data
from radex_emcee_b200.sampler import CudaEngine, SLEDModel, StretchSampler

ap = argparse.ArgumentParser()
ap.add_argument("--ncomp", type=int, default=2)
ap.add_argument("--log2w", type=int, default=14)
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--stop", default="pyradex")
ap.add_argument("--kernel", type=int, default=0, help="rb_opts.kernel (3 forces the fused single-launch lnprob)")
ap.add_argument("--spread", type=float, default=1e-3, help="sigma of the starting ball around p0")
ap.add_argument("--nsources", type=int, default=1, help="fit the first NSOURCES rows of flux.dat concurrently (config 4; ncomp 1)")
ap.add_argument("--no-native", action="store_true", help="per-half-step calls from Python instead of rb_stretch_run_dev")
ap.add_argument("--parity-split", action="store_true", help="randomize_split=False")
ap.add_argument("--spec", type=int, default=0, help="rb_opts.spec_half: 0 automatic, -1 never, 1 whenever the fused launch is used")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nw = args.walkers or (1 << args.log2w)
models = None
if args.ncomp == 1 and args.nsources > 1:
    data = read_data(ROOT + "/data/flux.dat")
    models, starts = [], []
    opts = _lib.default_opts(stop_rule=0 if args.stop == "pyradex" else 1, kernel=args.kernel, spec_half=args.spec)
    for k, nm in enumerate(list(data)[:args.nsources]):
        z, lw, jup, flux, eflux = get_source(nm, data)
        tbg, ra, bounds, p0 = er1.source_setup(z)
        models.append(SLEDModel(1, jup, flux, eflux, bounds, tbg, opts=opts))
        starts.append(p0 + args.spread * np.random.default_rng(20170914 + k).standard_normal((nw // args.nsources, 4)))
    pos = np.vstack(starts)
elif args.ncomp == 1:
    z, lw, jup, flux, eflux = get_source("G09v1.97", read_data(ROOT + "/data/flux.dat"))
    tbg, ra, bounds, p0 = er1.source_setup(z)
    T_d = None
else:
    z, T_d, lw, jup, flux, eflux = get_source("G09v1.97", read_data(ROOT + "/data/flux_for2p.dat"))
    tbg, ra, bounds, p0 = er2.source_setup(z)
    p0[3] += 0.1      # cold size > warm size so the whole starting ball has a finite prior
ctx = _lib.Context(_lib.MolData(os.path.join(ROOT, "radex_emcee_b200", "data", "co.dat")), lr)
opts = _lib.default_opts(stop_rule=0 if args.stop == "pyradex" else 1, kernel=args.kernel, spec_half=args.spec)
if models is None:
    eng = CudaEngine(ctx, SLEDModel(args.ncomp, jup, flux, eflux, bounds, tbg, T_d=T_d, opts=opts))
    pos = p0 + args.spread * np.random.default_rng(20170914).standard_normal((nw, 4 * args.ncomp))
else:
    eng = CudaEngine(ctx, models)
s = StretchSampler(nw, 4 * args.ncomp, eng, seed=1, nsources=args.nsources, native=not args.no_native,
                   randomize_split=not args.parity_split)
s.run_mcmc(pos, args.warmup, store=False)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
solves0 = int(eng.total_solves.item()) + s.total_solves
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
s.run_mcmc(None, args.steps, store=False)
e1.record()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1), float(int(eng.total_solves.item()) + s.total_solves - solves0)], dtype=torch.float64, device="cuda")
if world > 1:
    mx = ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = ms.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    t_ms, solves = float(mx[0]), float(sm[1])
else:
    t_ms, solves = float(ms[0]), float(ms[1])
acc = float(np.mean(s.acceptance_fraction))
if rank == 0:
    print(json.dumps({"metric": "walker-steps/s", "value": nw * args.steps / (t_ms * 1e-3), "n_gpus": world, "walkers": nw,
                      "ncomp": args.ncomp, "steps": args.steps, "ms_per_step": t_ms / args.steps, "solves_per_s": solves / (t_ms * 1e-3),
                      "solves_per_walker_step": solves / (nw * args.steps), "acceptance_fraction": acc, "wall_s": wall,
                      "stop": args.stop, "kernel": args.kernel, "launches": eng.launches, "nsources": args.nsources,
                      "native_loop": s.native, "randomize_split": not args.parity_split, "spec_half": args.spec}))
if world > 1:
    dist.destroy_process_group()
