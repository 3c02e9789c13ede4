"""Latency of small DeepZ calls through the blocking host-buffer API (BASELINE configs C1 / C2)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clr_b200
from clr_b200 import runtime, workloads
ctx = runtime.Context(0)
gold = os.path.join(ROOT, "tests", "golden", "nets")
tora = clr_b200.FeedforwardNetwork.load_npz(os.path.join(gold, "TORA_ReLU.npz"))
acc = clr_b200.FeedforwardNetwork.load_npz(os.path.join(gold, "ACC_relu.npz"))
dt, da = ctx.upload(tora), ctx.upload(acc)
g = clr_b200.split_box(*workloads.TORA_X0, [10, 10, 10, 1])
def bench(f, reps=200):
    for _ in range(10): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e6
us = bench(lambda: ctx.deepz_boxgrid(dt, g.partition, g.centers, g.radii, post=("shift", -10.0), stats=False))
print(f"C2 TORA 10^3 cells per call: {us:.1f} us/call = {1000/us:.2f} M propagations/s")
A = np.zeros((5, 6)); A[2, 4] = 1; A[3, 0], A[3, 3] = 1, -1; A[4, 1], A[4, 4] = 1, -1
d = np.array([30.0, 1.4, 0, 0, 0])
rng = np.random.default_rng(0)
C, off, G = workloads.random_zonotopes(rng, 1, 6, 14, 14, 0.0, 1.0, 0.05)
C += np.array([90.0, 32.0, 0.0, 10.0, 30.0, 0.0])
us = bench(lambda: ctx.deepz_zonotopes(da, C, off, G, pre=(A, d), stats=False))
print(f"C1 ACC single zonotope (n = 6, p = 14): {us:.1f} us/propagation")
ctx.close()

# C3: Quadrotor controller (12 -> 64^3 -> 3, sigmoid), 4^6 = 4096 cells of X0 (models/Quadrotor/Quadrotor.jl:98-99)
ctx = runtime.Context(0)
quad = clr_b200.FeedforwardNetwork.load_npz(os.path.join(gold, "Quadrotor.npz"))
dq = ctx.upload(quad)
r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
gq = clr_b200.split_box(-r, r, [4] * 6 + [1] * 6)
out = ctx.deepz_boxgrid(dq, gq.partition, gq.centers, gq.radii)
us = bench(lambda: ctx.deepz_boxgrid(dq, gq.partition, gq.centers, gq.radii, stats=False), reps=50)
fl = out["stats"]["flops"]
print(f"C3 Quadrotor 4096 cells per call: {us:.1f} us/call = {4096/us:.2f} M propagations/s, {fl/us/1e6:.2f} algorithmic TFLOP/s, "
      f"p_in/cell {[x/4096 for x in out['stats']['sum_p_in']]}")
ctx.close()
