#!/usr/bin/env python
"""bench.py -- LVG solves/s of the batched forward-model sweep (BASELINE.json configs[1]).

One "step" = one pass of the hot path over one batch: 2^20 random (Tkin, n_H2, N_CO/dv) CO LVG
solves (41 levels / 40 lines), drawn as SURVEY.md 8(d) config 2, each carried to the reference's
own stop rule (pyradex: sum|dx| < 1e-16 after > 10 iterations, cap 200) unless --stop radex.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2n 20] [--stop pyradex|radex]

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, kernel timed with CUDA events on
its launch stream, L2 flushed between timed steps.  `e2e`: same workload through the C ABI's
host-pointer entry (rb_solve_batch) with pinned host buffers, H2D and D2H inside the timed region.
`--impl reference`: the reference's CPU implementation of the same path (the bit-exact C
restatement in oracle/, since the reference ships only a macOS binary) on all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TBG = 2.7315 * (1 + 3)          # config 2: z = 3 background
# algorithmic flops: per matrix iteration and per solve prologue/epilogue, CO
F_ITER = 72368.0
F_PRO = 1.3e4
F_EPI = 4.0e3
# DRAM bytes of one 2^20 step (all launches), dram__bytes_read.sum + dram__bytes_write.sum of the ncu pass kept in
# profiles/r1e_launches.csv (IDs 34-44): launch A 8.65 GB (parks 1 KB of state per model and 10.9 KB of capture per
# cacheable model), B 0.18 GB, the five cached-engine launches 8.4 GB (read them back, write the results), C 0.3 GB
TRAFFIC_NCU_2P20 = 17.5e9


def draw(n, seed):
    rng = np.random.default_rng(seed)
    out = np.empty((0, 3))
    while out.shape[0] < n:
        m = int(1.6 * (n - out.shape[0])) + 16
        ln, lt, lN = rng.uniform(2, 7, m), rng.uniform(np.log10(TBG), 3, m), rng.uniform(15.5, 19.5, m)
        ok = (lN - ln > 10.0) & (lN - ln < 17.5)
        out = np.vstack([out, np.column_stack([10 ** lt, 10 ** ln, 10 ** lN])[ok]])
    out = out[:n]
    return np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1]), np.ascontiguousarray(out[:, 2])


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference(tk, nh2, cd, stop_rule, max_seconds, threads):
    """Time the oracle (bit-exact restatement of the reference's CPU path) on `threads` host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle
    molfile = os.path.join(ROOT, "radex_emcee_b200", "data", "co.dat")
    # calibrate on a few solves, then size the sample for ~max_seconds of wall time
    o = Oracle(molfile)
    t0 = time.perf_counter()
    o.solve_batch(tk[:24], 0.25 * nh2[:24], 0.75 * nh2[:24], cd[:24], tbg=TBG, stop_rule=stop_rule)
    per = (time.perf_counter() - t0) / 24
    n = int(min(tk.size, max(threads * 8, max_seconds / per * threads)))
    chunks = np.array_split(np.arange(n), threads)
    oracles = [Oracle(molfile) for _ in range(threads)]

    def work(i):
        idx = chunks[i]
        r = oracles[i].solve_batch(tk[idx], 0.25 * nh2[idx], 0.75 * nh2[idx], cd[idx], tbg=TBG, stop_rule=stop_rule)
        return int(r["niter"].sum())

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        iters = sum(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return n / dt, n, dt, iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--stop", default="pyradex", choices=["pyradex", "radex"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sort", default="", help="debug: order the models by cd | tk | top (highest thick line, from a first pass)")
    ap.add_argument("--keep", default="", help="debug: small | big | k57 | k8 -- keep only the models whose lead block (from a first pass) has <= 16 | > 16 | 20..28 | > 28 levels, tiled to n")
    ap.add_argument("--same", type=int, default=-1, help="debug: every model is a copy of draw #SAME (I-cache experiments)")
    ap.add_argument("--kernel", type=int, default=0, help="rb_opts.kernel: 0 default, 1 v1 LU, 2 v2 without caching, 3 single launch, 4 without the half-warp engine")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = 1 << args.log2n
    stop_rule = 0 if args.stop == "pyradex" else 1
    workload = ("batched forward-model sweep: 2^%d random (Tkin,n_H2,N_CO/dv) CO LVG solves, 41 levels, "
                "tbg=%.3f K, stop=%s" % (args.log2n, TBG, args.stop))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        tk, nh2, cd = draw(n, 0)
        vals = []
        per_step = max(5.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
        sample = 0
        for s in range(args.warmup + args.steps):
            v, sample, dt, _ = cpu_reference(tk, nh2, cd, stop_rule, per_step, cores)
            if s >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals]))
        line = {"impl": "reference", "metric": "LVG solves/s", "value": value, "unit": "solves/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([d for _, d in vals]) * 1e3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                                 "sample": "first %d draws of the sweep per step, %d threads; oracle/ is the bit-exact C "
                                           "restatement of the reference (its Fortran ships only as a macOS binary)" % (sample, cores)},
                "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from radex_emcee_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    mol = _lib.MolData(os.path.join(ROOT, "radex_emcee_b200", "data", "co.dat"))
    ctx = _lib.Context(mol, local_rank)
    L = _lib.load()
    nl, nn, npart = mol.nlev, mol.nline, mol.npart

    tk, nh2, cd = draw(n, rank)                      # weak scaling: every rank gets its own 2^k draws
    if args.same >= 0:
        tk[:], nh2[:], cd[:] = tk[args.same], nh2[args.same], cd[args.same]
    if args.sort in ("cd", "tk"):
        o = np.argsort(cd if args.sort == "cd" else tk, kind="stable")
        tk, nh2, cd = tk[o].copy(), nh2[o].copy(), cd[o].copy()
    dens = np.zeros((n, npart))
    for p, pid in enumerate(mol.partner_id):
        dens[:, p] = {2: 0.25, 3: 0.75}.get(int(pid), 0.0) * nh2
    d_tk, d_dens, d_cd = (torch.from_numpy(a).to(dev) for a in (tk, dens, cd))
    d_x = torch.empty((n, nl), dtype=torch.float64, device=dev)
    d_tex = torch.empty((n, nn), dtype=torch.float64, device=dev)
    d_tau = torch.empty_like(d_tex)
    d_surf = torch.empty_like(d_tex)
    d_it = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    opts = _lib.default_opts(stop_rule=stop_rule, kernel=args.kernel)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    def launch():
        _lib.check(L.rb_solve_batch_dev(ctx.handle, n, d_tk.data_ptr(), d_dens.data_ptr(), d_cd.data_ptr(), 1.0, TBG, 2,
                                        C.byref(opts), d_x.data_ptr(), d_tex.data_ptr(), d_tau.data_ptr(),
                                        d_surf.data_ptr(), d_it.data_ptr(), d_st.data_ptr()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.sort == "top":      # oracle ordering: by the highest optically thick line of the converged model
        launch()
        torch.cuda.synchronize(dev)
        tau = d_tau.cpu().numpy()
        thick = np.abs(tau) * 0.5 >= 0.01
        top = np.where(thick.any(axis=1), thick.shape[1] - 1 - np.argmax(thick[:, ::-1], axis=1), -1)
        o = np.argsort(top, kind="stable")
        tk, nh2, cd, dens = tk[o].copy(), nh2[o].copy(), cd[o].copy(), dens[o].copy()
        d_tk.copy_(torch.from_numpy(tk)); d_cd.copy_(torch.from_numpy(cd)); d_dens.copy_(torch.from_numpy(dens))
    if args.keep in ("small", "big", "k57", "k8"):
        launch()
        torch.cuda.synchronize(dev)
        tau = d_tau.cpu().numpy()
        thick = ~(np.abs(tau) * 0.5 < np.float32(0.01))
        top = np.where(thick.any(axis=1), thick.shape[1] - np.argmax(thick[:, ::-1], axis=1), -1)   # upper level of the highest thick line
        key = np.maximum(3, (top + 1 + 4) >> 2)
        sel = np.nonzero({"small": key <= 4, "big": key > 4, "k57": (key > 4) & (key < 8), "k8": key >= 8}[args.keep])[0]
        o = np.resize(sel, n)
        tk, nh2, cd, dens = tk[o].copy(), nh2[o].copy(), cd[o].copy(), dens[o].copy()
        d_tk.copy_(torch.from_numpy(tk)); d_cd.copy_(torch.from_numpy(cd)); d_dens.copy_(torch.from_numpy(dens))
    for _ in range(args.warmup):
        launch()
    torch.cuda.synchronize(dev)
    launches_before = ctx.counters()[1]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- kernel-only: K steps, each bracketed by events on the launch stream, L2 flushed between ----
    barrier()
    t_wall0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        launch()
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = float(sum(step_ms))
    total_iters, launches_after = ctx.counters()      # iterations of the last step (device-counted)
    launches_timed = launches_after - launches_before  # kernels launched inside the kernel-only timed region
    cache_stats = ctx.cache_stats()                   # (cached iterations, captures, invalidations), last launch
    niter_host = d_it.cpu().numpy()
    status_host = d_st.cpu().numpy()

    # ---- e2e: host buffers through the C ABI (pinned), H2D + D2H inside the timed region --------------
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    h_tk, h_cd = pinned(n, torch.float64), pinned(n, torch.float64)
    h_dens = pinned((n, npart), torch.float64)
    h_tk.copy_(torch.from_numpy(tk)); h_cd.copy_(torch.from_numpy(cd)); h_dens.copy_(torch.from_numpy(dens))
    h_x, h_tex, h_tau, h_surf = pinned((n, nl), torch.float64), pinned((n, nn), torch.float64), \
        pinned((n, nn), torch.float64), pinned((n, nn), torch.float64)
    h_it, h_st = pinned(n, torch.int32), pinned(n, torch.int32)

    def e2e_call():
        _lib.check(L.rb_solve_batch(ctx.handle, n, h_tk.data_ptr(), h_dens.data_ptr(), h_cd.data_ptr(), 1.0, TBG, 2,
                                    C.byref(opts), h_x.data_ptr(), h_tex.data_ptr(), h_tau.data_ptr(),
                                    h_surf.data_ptr(), h_it.data_ptr(), h_st.data_ptr()))

    e2e_call()                                        # warm-up (allocates the ctx scratch)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()                                    # synchronous: returns after the D2H copies
    barrier()
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=5)
    h2d = (2 * n + n * npart) * 8
    d2h = (n * nl + 3 * n * nn) * 8 + 2 * n * 4

    # ---- max over ranks ------------------------------------------------------------------------------------
    t = torch.tensor([kern_ms, e2e_s * 1e3, float(total_iters)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        kern_ms, e2e_ms, iters_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        kern_ms, e2e_ms, iters_all = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        peaks, peak_src = load_peaks()
        fp64_peak = ctx.fp64_peak_tflops()
        ms_per_step = kern_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        e2e_value = world * n / (e2e_ms / args.steps * 1e-3)
        flops_launch = iters_all / world * F_ITER + n * (F_PRO + F_EPI)          # per launch (one rank)
        achieved = flops_launch / (ms_per_step * 1e-3) * 1e-12
        bytes_launch = h2d + d2h                                                  # algorithmic HBM bytes
        line = {
            "metric": "LVG solves/s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "models_per_gpu": n, "l2": "256 MiB flush write between timed steps",
                       "outputs": "xpop,tex,tau,surf,niter,status",
                       "kernel": {0: "k_lvg_solve_v2 launches A, B, C (frozen-top caching, ordered by lead-block size) + k_lvg_small<3..7> (cached engines per lead-block size, two models per warp up to 16 levels)",
                                  1: "k_lvg_solve_v1", 2: "k_lvg_solve_v2 (no caching)",
                                  3: "k_lvg_solve_v2 (frozen-top caching, single launch)",
                                  4: "k_lvg_solve_v2 (frozen-top caching, two launches ordered by lead-block size)"}[args.kernel]},
            "iters_per_solve": iters_all / (world * n),
            "matrix_iterations_per_s": iters_all / (ms_per_step * 1e-3),
            "frac_iterations_cached": cache_stats[0] / max(1, total_iters),
            "captures_per_solve": cache_stats[1] / n, "invalidations_per_solve": cache_stats[2] / n,
            "frac_at_maxiter": float((status_host & 4).astype(bool).mean()),
            "frac_nonfinite": float((status_host & 8).astype(bool).mean()),
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_timed,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak,
                         "traffic": TRAFFIC_NCU_2P20 if (args.log2n == 20 and args.kernel == 0 and stop_rule == 0
                                                          and not args.keep and not args.sort and args.same < 0) else None,
                         "traffic_unit": "bytes per step, all launches (ncu, profiles/r1e_launches.csv); algorithmic bytes = h2d + d2h",
                         "peak_source": "rb_fp64_peak DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 "
                                        "figure); nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2",
                         "flops_per_iter": F_ITER,
                         "hbm": {"achieved_gbs": bytes_launch / (ms_per_step * 1e-3) * 1e-9,
                                 "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src}},
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu and world == 1:
            v, sample, dt, _ = cpu_reference(tk, nh2, cd, stop_rule, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": 1, "kind": "port",
                                    "sample": "first %d draws of the same sweep, 1 thread, %.1f s; oracle/ is the "
                                              "bit-exact C restatement of the reference" % (sample, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
