"""Times the batched _project_oa kernels (clr_project_oa) with device-resident arrays:
   python profiles/project_bench.py [N=200000]   -- n = 6 variables, order_t = 8, order_q = 3 (84 monomials), 4 kept."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import runtime
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
n, oT, oQ, vk = 6, 8, 3, [0, 1, 2, 3]
import itertools
def table(n, order):
    out = []
    for d in range(order + 1):
        def rec(prefix, left, k):
            if k == n - 1:
                out.append(prefix + [left]); return
            for e in range(left, -1, -1):
                rec(prefix + [e], left - e, k + 1)
        rec([], d, 0)
    return np.array(out, dtype=np.int32)
exps = table(n, oQ); M = exps.shape[0]
ctx = runtime.Context(0); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
coeffs = torch.randn((N, n, oT + 1, M), dtype=torch.float64, device=dev, generator=g) * 0.01
rem = torch.sort(torch.randn((N, n, 2), dtype=torch.float64, device=dev, generator=g) * 1e-6, dim=2).values.contiguous()
dt = torch.full((N,), 0.1, dtype=torch.float64, device=dev)
c = torch.empty((N, len(vk)), dtype=torch.float64, device=dev)
G = torch.empty((N, 2 * n, len(vk)), dtype=torch.float64, device=dev)
nc = torch.empty((N,), dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
run = lambda: ctx.project_oa_dev(n, oT, exps, N, coeffs, rem, dt, vk, True, c, G, nc)
run(); ctx.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(5):
    flush.fill_(1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        a.record(stream); run(); b.record(stream)
    ctx.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
alg = N * (8 * len(vk) * (oT + 1) * M + 16 * len(vk) + 8 + 8 * len(vk) * (1 + 2 * n) + 4)
print(f"project_oa: {ms:.3f} ms for {N} reach-sets ({coeffs.numel() * 8 / 1e9:.2f} GB of coefficients, {len(vk)} of {n} components kept): "
      f"{N / ms * 1e3:.3e} reach-sets/s, algorithmic {alg / ms / 1e6:.0f} GB/s")
ctx.close()
