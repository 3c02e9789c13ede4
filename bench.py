#!/usr/bin/env python
"""bench.py -- the hot path of ClosedLoopReachability.jl on B200: zonotope controller propagations/s.

A "step" is one pass of the hot path over one batch of synthetic input: every split cell of the
C4 scaling configuration (BASELINE.json configs[3]: 10^6 box cells through a 6-input 4x100 ReLU
network, Float64) goes through controller_forward with DeepZ on the GPU.  Cells are sharded over the
ranks by contiguous cell-id range (weak scaling: 10^6 cells per GPU); there is no collective on the
data path, only the final gather of the per-GPU hulls.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (C restatement of the reference)

One JSON line is printed by rank 0.  `value` = cells x steps / device time with inputs resident in HBM
(CUDA events on the library's stream, L2 flushed between steps, max over ranks); `e2e` = the same
metric through the host-buffer API call (H2D of the tables, D2H of all boxes inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DFMA_PEAK_TFLOPS = 36.2   # measured FP64 FMA issue peak, same file (the rollout kernel is FP64-issue bound, not HBM bound)
DMMA_PEAK_TFLOPS = 37.0   # measured on this pool's B200: profiles/microbench/r01_dmma_bench_b200.txt (mma.sync f64)
METRIC = "zonotope controller propagations/sec (cells x steps)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regime", default=os.environ.get("CLR_BENCH_REGIME", "stable"), choices=["stable", "mixed", "dense"])
    ap.add_argument("--cells-per-dim", type=int, default=10)
    ap.add_argument("--no-secondary", action="store_true", help="skip the simulate() rollout line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traj", type=int, default=1000000)
    ap.add_argument("--no-regimes", action="store_true", help="skip the mixed / dense regime lines")
    ap.add_argument("--shards", default="cyclic", choices=["cyclic", "equal", "cost"],
                    help="N > 1: block-cyclic shards (default), contiguous equal counts, or contiguous cost-weighted ranges")
    return ap.parse_args()


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line (the driver's contract).  Libraries write there too (NCCL prints its version
    banner on stdout when NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, world):
    """C4: box [c0-s, c0+s]^6 split [10]^5 x [10 N]: with N ranks the SAME box is split N times finer in the last (slowest)
    dimension -- the way a verification run that needs more precision grows -- so the job has N x 10^6 cells of about
    the same cost each (round 1 made the box N times longer instead: its far end is costlier per cell, which put
    8 % of 'scaling loss' into the workload itself)."""
    from clr_b200 import workloads
    from clr_b200.splitters import split_box
    net = workloads.synthetic_relu_net()
    c0, s = {"dense": (0.0, 1.0), "mixed": (0.3, 0.1), "stable": (0.3, 0.01)}[args.regime]
    k = args.cells_per_dim
    grid = split_box([c0 - s] * 6, [c0 + s] * 6, [k] * 5 + [k * world])
    return net, grid


def cpu_baseline(net, grid, budget_s=12.0):
    """The C restatement of the reference (oracle/, OpenMP over cells, all host cores) on a bounded sample."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import cbind
    on = cbind.Net(net.as_tuples())
    cores = cbind.num_threads()
    n, chunk, done, t0 = len(grid), 4000, 0, time.perf_counter()
    cbind.deepz_boxgrid(on, grid.partition, grid.centers, grid.radii, cell_begin=0, cell_count=256)   # warm-up
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s and done < n:
        cnt = min(chunk, n - done)
        cbind.deepz_boxgrid(on, grid.partition, grid.centers, grid.radii, cell_begin=done, cell_count=cnt)
        done += cnt
        chunk = min(chunk * 2, 64000)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "propagations/s", "cores": cores, "kind": "port",
            "sample": f"first {done} cells of the same split, {dt:.1f} s, C oracle (OpenMP, -O3)"}


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun pins it to 1; the CPU arm uses every host core
    net, grid = workload(args, 1)
    from oracle import cbind
    on = cbind.Net(net.as_tuples())
    cores = cbind.num_threads()
    per_step = {"stable": 100000, "mixed": 30000, "dense": 6000}[args.regime]
    for _ in range(max(args.warmup, 1)):
        cbind.deepz_boxgrid(on, grid.partition, grid.centers, grid.radii, cell_begin=0, cell_count=per_step // 10)
    t0 = time.perf_counter()
    for s in range(args.steps):
        b = (s * per_step) % (len(grid) - per_step)
        cbind.deepz_boxgrid(on, grid.partition, grid.centers, grid.radii, cell_begin=b, cell_count=per_step)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "propagations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4 synthetic scaling, 6->100x4->1 ReLU, regime={args.regime}; bounded sample of "
                                   f"{per_step} cells per step of the 10^6-cell split"},
            "cpu_baseline": {"value": v, "unit": "propagations/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} cells per step x {args.steps} steps; the reference is Julia "
                                       "(absent from this image), so its C restatement in oracle/ is timed"},
            "e2e": {"value": v, "unit": "propagations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def hw_utilisation(regime):
    """Hardware counters of the main-pass launch of this workload from the committed ncu --set full capture (not measured
    in this run: ncu replays a kernel ~40 times and a number taken under a profiler is never a bench value)."""
    for rel in (os.path.join("profiles", "r02", f"ncu_deepz_{regime}.json"), os.path.join("profiles", "r02", "ncu_deepz_stable.json"),
                os.path.join("profiles", "r01", "ncu_deepz_bench_launch.json")):
        try:
            with open(os.path.join(ROOT, rel)) as fh:
                d = json.load(fh)
            if regime not in rel and regime != "stable":
                continue
            return {"dmma_pipe_active_pct": d.get("dmma_pipe_active_pct"), "issue_active_pct": d.get("issue_active_pct"),
                    "fp64_pipe_pct": d.get("fp64_pipe_pct"), "traffic_bytes_per_launch": d.get("traffic"),
                    "cells_in_capture": d.get("cells"), "source": f"committed ncu capture {rel} (not measured in this run)"}
        except Exception:
            continue
    return None


def measure_regime(args, regime, steps, warmup, ctx, dev, stream, flush, rank, world, fp64_peak, e2e_steps):
    """One regime of C4 on this rank's shard: device-resident throughput (CUDA events on the library's stream, L2 flushed
    between steps), per-pass times, algorithmic FLOPs, and the end-to-end number through the host-buffer API."""
    import torch
    import torch.distributed as dist
    from clr_b200 import _lib, parallel, runtime
    a2 = argparse.Namespace(**vars(args))
    a2.regime = regime
    net, grid = workload(a2, world)
    dn = ctx.upload(net)
    m = 1
    pp = runtime.PrePostArg(None, None)
    cc = torch.from_numpy(np.concatenate(grid.centers)).to(dev)
    rr = torch.from_numpy(np.concatenate(grid.radii)).to(dev)
    # shards: block-cyclic by default (rank r takes the blocks r, r + N, ... of 1024 cells: every share is spread over the
    # whole split, so the ranks take the same time without a cost model); --shards equal / cost give contiguous ranges
    cost, blk, stride = None, 0, 0
    if world > 1 and args.shards == "cyclic":
        blk, stride = 1024, 1024 * world
        nb = (len(grid) + blk - 1) // blk
        cnt = sum(min(blk, len(grid) - j * blk) for j in range(rank, nb, world))
        b = rank * blk
    else:
        if world > 1 and args.shards == "cost":
            bounds, cost = parallel.balanced_bounds(ctx, dn, grid, world)
        else:
            bounds = [parallel.shard_range(len(grid), r, world)[0] for r in range(world)] + [len(grid)]
        b, cnt = bounds[rank], bounds[rank + 1] - bounds[rank]
    d_lo = torch.empty((cnt, m), dtype=torch.float64, device=dev)
    d_hi = torch.empty((cnt, m), dtype=torch.float64, device=dev)
    d_hl = torch.empty(m, dtype=torch.float64, device=dev)
    d_hh = torch.empty(m, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_dev():
        ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, b, cnt, pp, d_lo, d_hi, d_hl, d_hh, block=blk, stride=stride)

    # algorithmic FLOPs of one step (statistics run, untimed)
    st = _lib.Stats()
    ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, b, cnt, pp, d_lo, d_hi, d_hl, d_hh, stats=st, block=blk, stride=stride)
    ctx.synchronize()
    flops_step = st.flops
    p_in = [st.sum_p_in[l] / cnt for l in range(5)]

    ctx.set_profiling(True)
    for _ in range(warmup):
        step_dev()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = ctx.launch_count
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    pass_ms = np.zeros(3)
    retried = big = 0
    wall0 = time.perf_counter()
    for s in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # L2 flush between timed iterations (not timed)
            ev0[s].record(stream)
            step_dev()
            ev1[s].record(stream)
        info = ctx.last_pass_info()
        pass_ms += np.array(info["pass_ms"])
        retried, big = info["retried"], info["big"]
    ctx.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(bb) for a, bb in zip(ev0, ev1))
    if world > 1:
        dist.barrier()
    rank_ms = parallel.gather_scalars(dev_ms / steps, device=dev)        # per-rank ms per step
    rank_flops = parallel.gather_scalars(flops_step, device=dev)
    rank_cells = parallel.gather_scalars(cnt, device=dev)
    t_ms = max(rank_ms) * steps
    total_cells = sum(rank_cells)
    value = total_cells * steps / (t_ms * 1e-3)
    hl, hh = parallel.gather_hull(d_hl, d_hh)   # final gather of the per-GPU hulls over NCCL (the only communication)
    hull = [float(hl[0]), float(hh[0])]

    # ---- end to end: host buffers through the public API (blocking call) -------------------------------
    out_lo = torch.empty((cnt, m), dtype=torch.float64).pin_memory().numpy()
    out_hi = torch.empty((cnt, m), dtype=torch.float64).pin_memory().numpy()
    ctx.deepz_boxgrid(dn, grid.partition, grid.centers, grid.radii, cell_begin=b, cell_count=cnt, stats=False,
                      out_lo=out_lo, out_hi=out_hi, block=blk, stride=stride)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.deepz_boxgrid(dn, grid.partition, grid.centers, grid.radii, cell_begin=b, cell_count=cnt, stats=False,
                          out_lo=out_lo, out_hi=out_hi, block=blk, stride=stride)
    torch.cuda.synchronize()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, device=dev)
    tab_bytes = 2 * 8 * int(np.sum(grid.partition))
    e2e = {"value": total_cells * e2e_steps / e2e_s, "unit": "propagations/s", "steps": e2e_steps,
           "h2d_bytes_per_step": tab_bytes, "d2h_bytes_per_step": int(2 * 8 * m * cnt),
           "note": "clr_deepz_boxgrid with host buffers: H2D of the split tables, D2H of every cell's box, blocking"}
    assert np.array_equal(out_lo, d_lo.cpu().numpy())

    # ---- roofline: WHOLE step (every pass, the pilot and the host gaps between them), per GPU --------------------------
    ms_step = t_ms / steps
    flops_per_gpu = sum(rank_flops) / world
    ach = flops_per_gpu / (ms_step * 1e-3) / 1e12
    dom = int(np.argmax(pass_ms))
    dom_ms = pass_ms[dom] / steps
    hw = hw_utilisation(regime)
    roofline = {"bound": "tensor", "achieved": ach, "peak": fp64_peak["value"], "unit": "TFLOP/s", "frac": ach / fp64_peak["value"],
                "traffic": hw["traffic_bytes_per_launch"] if hw else None,
                "traffic_source": hw["source"] if hw else None,
                "what": "achieved = algorithmic FLOPs of one step (per GPU: 2 m n (p + 1) per layer and cell, p counted by the "
                        "kernel's statistics run) / ms_per_step -- the whole step, all passes",
                "peak_source": fp64_peak["source"],
                "algorithmic_flops_per_step": flops_per_gpu,
                "dominant_kernel": {"name": ["deepz_kernel (main pass, smem store)", "deepz_kernel (retry pass)",
                                             "deepz_kernel (big-cell pass, global store)"][dom],
                                    "ms_per_step": dom_ms, "share_of_step": dom_ms / ms_step,
                                    "algorithmic_tflops_if_it_did_all_the_work": flops_step / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else None},
                "pass_ms_per_step": (pass_ms / steps).tolist(), "mean_generators_in": p_in,
                "cells_retried": int(retried), "cells_big": int(big),
                "hardware_utilisation": hw}
    return {"value": value, "ms_per_step": ms_step, "rank_ms_per_step": rank_ms, "rank_cells": [int(c) for c in rank_cells],
            "rank_algorithmic_flops": rank_flops, "shard_cost_model": cost, "clocks": clocks, "e2e": e2e, "launches": int(launches),
            "roofline": roofline, "wall": wall, "hull": hull, "cnt": cnt, "net": net, "grid": grid, "steps": steps, "warmup": warmup}


def run_ours(args):
    import torch
    import clr_b200
    from clr_b200 import parallel, runtime, workloads
    rank, local, world = dist_env()
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = runtime.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # FP64 tensor roofline denominator, measured on this device in this run (MEASURED_PEAKS.json has no FP64 entry)
    live = ctx.measure_fp64_peak(30.0)
    fp64_peak = {"value": live, "source": f"FP64 mma.sync (DMMA) issue peak measured in this run on this GPU (clr_measure_fp64_peak, "
                                           f"{live:.2f} TFLOP/s); committed microbenchmark: {DMMA_PEAK_TFLOPS} (profiles/microbench/"
                                           "r01_dmma_bench_b200.txt; cuBLAS DGEMM 8192^3 = 35.5); MEASURED_PEAKS.json has no FP64 entry"}
    e2e_steps = max(3, min(args.steps, 10))
    main = measure_regime(args, args.regime, args.steps, args.warmup, ctx, dev, stream, flush, rank, world, fp64_peak, e2e_steps)
    net, grid, cnt = main["net"], main["grid"], main["cnt"]

    line = {"metric": METRIC, "value": main["value"], "unit": "propagations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4 synthetic scaling: {len(grid) // world} box cells per GPU (split [10]^5 x [10 N] of a 6-dim box), "
                                   f"6->100x4->1 ReLU network (seed 4), Float64 DeepZ, regime={args.regime}",
                       "cells_per_gpu": len(grid) // world, "regime": args.regime, "l2": "256 MB buffer written between timed steps",
                       "hull": main["hull"],
                       "shards": "one shard" if world == 1 else
                                 {"cyclic": "block-cyclic: rank r takes the blocks r, r + N, ... of 1024 cells",
                                  "equal": "contiguous cell-id ranges of equal counts",
                                  "cost": "contiguous cell-id ranges, boundaries weighted by a pilot's per-cell cost"}[args.shards]},
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["launches"], "roofline": main["roofline"],
            "wall_s_timed_region": main["wall"],
            "per_rank": {"ms_per_step": main["rank_ms_per_step"], "cells": main["rank_cells"],
                         "algorithmic_flops_per_step": main["rank_algorithmic_flops"], "shard_cost_model": main["shard_cost_model"]}}

    # ---- the other regimes of SURVEY 8d (the box as is = "dense", and the intermediate one), same split, fewer steps -----
    if not args.no_regimes:
        line["regimes"] = {}
        for reg in ("stable", "mixed", "dense"):
            if reg == args.regime:
                continue
            r = measure_regime(args, reg, 3, 3, ctx, dev, stream, flush, rank, world, fp64_peak, 3)
            line["regimes"][reg] = {"value": r["value"], "unit": "propagations/s", "ms_per_step": r["ms_per_step"], "steps": 3,
                                    "warmup": 3, "e2e": r["e2e"]["value"], "rank_ms_per_step": r["rank_ms_per_step"],
                                    "roofline": {k: r["roofline"][k] for k in ("achieved", "peak", "frac", "algorithmic_flops_per_step",
                                                                                "pass_ms_per_step", "mean_generators_in", "cells_retried",
                                                                                "cells_big", "hardware_utilisation")},
                                    "clocks": r["clocks"]}

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(net, grid)
    elif rank == 0:
        line["cpu_baseline"] = None

    # ---- secondary metric: simulate() rollouts (Unicycle, 10^6 trajectories x 50 periods) -----------------
    if not args.no_secondary:
        uni = clr_b200.FeedforwardNetwork.load_npz(os.path.join(ROOT, "tests", "golden", "nets", "Unicycle.npz"))
        du = ctx.upload(uni)
        N, iters = args.traj, 50
        rng = np.random.default_rng(5 + rank)
        x0, Wd = workloads.unicycle_samples(rng, N, iters)
        d_x0 = torch.from_numpy(x0).to(dev)
        d_Wd = torch.from_numpy(Wd).to(dev)
        d_X = torch.empty((iters, N, 4), dtype=torch.float64, device=dev)
        d_U = torch.empty((iters, N, 2), dtype=torch.float64, device=dev)
        ppu = runtime.PrePostArg(None, ("shift", -20.0))
        res = {}
        for alg in ("Tsit5", "RK4"):
            def roll():
                ctx.rollout_dev(du, "Unicycle", N, iters, d_x0, d_Wd, 0.0, 0.2, 10.0, alg, 1e-6, 1e-3, 4, 1, ppu, d_X, d_U)
            roll()
            ctx.synchronize()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            with torch.cuda.stream(stream):
                a.record(stream)
                for _ in range(reps):
                    roll()
                c.record(stream)
            ctx.synchronize()
            ms = parallel.max_over_ranks(a.elapsed_time(c) / reps, device=dev)
            byts = 88.0 * N * iters                      # 88 B per trajectory-period
            flops = 6000.0 * N * iters                   # controller: 2 * (500*4 + 2*500) per trajectory-period
            res[alg] = {"value": world * N / (ms * 1e-3), "unit": "trajectories/s", "ms": ms,
                        "hbm_GBps_algorithmic": byts / (ms * 1e-3) / 1e9,
                        "controller_fp64_TFLOPs": flops / (ms * 1e-3) / 1e12,
                        "frac_of_fp64_fma_peak": flops / (ms * 1e-3) / 1e12 / DFMA_PEAK_TFLOPS}
        line["secondary"] = {"metric": "sim trajectories/sec", "config": f"Unicycle, {N} trajectories per GPU x {iters} "
                             "control periods, 4->500->2 ReLU controller; Tsit5 = the reference's solver, RK4 = 4 fixed steps / period",
                             "kernels": "per control period one controller_cw_kernel launch (packed controller in the constant bank, "
                                        "weights in uniform registers) + one rollout_ode_kernel launch",
                             **res}
        # end to end through the host-buffer call (x0 / disturbances up, X_end / U back, blocking) on a bounded sample, and
        # the CPU restatement (OpenMP over trajectories) on the host cores next to it -- rank 0, one GPU only
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            ns = min(N, 100000)
            # pinned host buffers (the contract's end-to-end figure); `pageable` = fresh NumPy arrays, what a caller that does not
            # page-lock its arrays sees (clr_host_register pins them in place)
            hx = torch.from_numpy(x0[:ns].copy()).pin_memory().numpy()
            hw = torch.from_numpy(np.ascontiguousarray(Wd[:, :ns])).pin_memory().numpy()
            hout = {"X_end": torch.empty((iters, ns, 4), dtype=torch.float64).pin_memory().numpy(),
                    "U": torch.empty((iters, ns, 2), dtype=torch.float64).pin_memory().numpy()}
            best = {}
            for tag, kw, xa, wa in (("pinned", {"out": hout}, hx, hw), ("pageable", {}, x0[:ns], Wd[:, :ns])):
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ctx.rollout(du, "Unicycle", xa, wa, 0.0, 0.2, 10.0, iters, post=("shift", -20.0), counts=False, **kw)
                    ts.append(time.perf_counter() - t0)
                best[tag] = ns / min(ts[1:])
            line["secondary"]["e2e"] = {"value": best["pinned"], "unit": "trajectories/s", "sample": f"{ns} trajectories, pinned host buffers",
                                        "pageable": best["pageable"], "h2d_bytes": int(8 * ns * (4 + iters)), "d2h_bytes": int(8 * ns * iters * 6)}
            from oracle import cbind
            nc = min(N, 20000)
            on = cbind.Net(uni.as_tuples())
            t0 = time.perf_counter()
            cbind.rollout(on, "Unicycle", 4, 1, 2, x0[:nc], Wd[:, :nc], 0.0, 0.2, 10.0, iters, post=("shift", -20.0))
            dt = time.perf_counter() - t0
            line["secondary"]["cpu_baseline"] = {"value": nc / dt, "unit": "trajectories/s", "cores": cbind.num_threads(),
                                                 "kind": "port", "sample": f"{nc} trajectories x {iters} periods, Tsit5"}
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
