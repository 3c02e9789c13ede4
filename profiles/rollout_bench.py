"""Times the simulate() rollout kernel (Unicycle, N trajectories x 50 periods) with device-resident I/O."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import runtime, workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
iters = 50
ctx = runtime.Context(0)
dev = torch.device("cuda", 0)
uni = clr_b200.FeedforwardNetwork.load_npz(os.path.join(ROOT, "tests", "golden", "nets", "Unicycle.npz"))
du = ctx.upload(uni)
x0, Wd = workloads.unicycle_samples(np.random.default_rng(5), N, iters)
d_x0 = torch.from_numpy(x0).to(dev); d_Wd = torch.from_numpy(Wd).to(dev)
d_X = torch.empty((iters, N, 4), dtype=torch.float64, device=dev)
d_U = torch.empty((iters, N, 2), dtype=torch.float64, device=dev)
d_na = torch.empty((iters, N), dtype=torch.int32, device=dev)
d_nr = torch.empty((iters, N), dtype=torch.int32, device=dev)
pp = runtime.PrePostArg(None, ("shift", -20.0))
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for alg, sub in (("Tsit5", 4), ("RK4", 4), ("RK4", 1)):
    def roll():
        ctx.rollout_dev(du, "Unicycle", N, iters, d_x0, d_Wd, 0.0, 0.2, 10.0, alg, 1e-6, 1e-3, sub, 1, pp, d_X, d_U, d_na, d_nr)
    roll(); ctx.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        a.record(stream); roll(); roll(); c.record(stream)
    ctx.synchronize()
    ms = a.elapsed_time(c) / 2
    print(alg, sub, f"{ms:.2f} ms  {N/ms*1e3:.3e} traj/s  mean accepted steps/period {d_na.double().mean().item():.2f} rejected {d_nr.double().mean().item():.3f}")
ctx.close()
