"""Times reduce_order_kernel (Girard order-2 reduction of the kept output zonotopes) on the Quadrotor controller:
   python profiles/reduce_bench.py [cells_per_dim=6]  -> cells = k^6 box cells, 198 generators x 3 rows each."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import runtime
k = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = runtime.Context(0)
dev = torch.device("cuda", 0)
net = clr_b200.FeedforwardNetwork.load_npz(os.path.join(ROOT, "tests", "golden", "nets", "Quadrotor.npz"))
dn = ctx.upload(net)
r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
g = clr_b200.split_box(-r, r, [k] * 6 + [1] * 6)
N, m, cap = k ** 6, 3, 200
cc = torch.from_numpy(np.concatenate(g.centers)).to(dev); rr = torch.from_numpy(np.concatenate(g.radii)).to(dev)
lo = torch.empty((N, m), dtype=torch.float64, device=dev); hi = torch.empty_like(lo)
pp = runtime.PrePostArg(None, None)
ctx.set_profiling(True)
ctx.deepz_boxgrid_dev(dn, 12, g.partition, cc, rr, 0, N, pp, lo, hi, keep_cap=cap)
ctx.synchronize()
print("deepz keep run:", ctx.last_pass_info())
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for order in (2, 1):
    d_c = torch.empty((N, m), dtype=torch.float64, device=dev)
    d_G = torch.empty((N, m * order, m), dtype=torch.float64, device=dev)
    ctx.fetch_reduced_dev(order, d_c, d_G); ctx.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for _ in range(5):
        flush.fill_(1); torch.cuda.synchronize()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream); ctx.fetch_reduced_dev(order, d_c, d_G); c.record(stream)
        ctx.synchronize()
        times.append(a.elapsed_time(c))
    ms = float(np.median(times))
    p = 198
    alg_bytes = N * 8 * (m * (1 + p) + m * (1 + m * order)) + 4 * N
    print(f"order {order}: {ms:.3f} ms for {N} cells  {N / ms * 1e3:.3e} zonotopes/s  algorithmic {alg_bytes / ms / 1e6:.1f} GB/s")
ctx.close()
