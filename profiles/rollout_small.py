"""Blocking clr_rollout latency against the number of trajectories (Unicycle, 50 periods, Tsit5): the fused kernel (one launch for the
whole horizon) against the split path (two launches per control period).  python profiles/rollout_small.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clr_b200
from clr_b200 import runtime, workloads
uni = clr_b200.FeedforwardNetwork.load_npz(os.path.join(ROOT, "tests", "golden", "nets", "Unicycle.npz"))
for split in ("0", "1"):
    os.environ["CLR_B200_RO_SPLIT"] = split
    os.environ["CLR_B200_RO_SPLIT_MIN"] = "0"
    ctx = runtime.Context(0)
    du = ctx.upload(uni)
    for N in (10, 100, 1000, 4096, 16384, 65536, 262144):
        x0, Wd = workloads.unicycle_samples(np.random.default_rng(5), N, 50)
        ctx.rollout(du, "Unicycle", x0, Wd, 0.0, 0.2, 10.0, 50, post=("shift", -20.0), counts=False)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.rollout(du, "Unicycle", x0, Wd, 0.0, 0.2, 10.0, 50, post=("shift", -20.0), counts=False)
        dt = (time.perf_counter() - t0) / reps
        print(f"split={split} N={N:7d}: {dt * 1e3:8.3f} ms per call")
    ctx.close()
