"""Small driver for ncu: one device-resident DeepZ call on a slice of the C4 workload.

    python profiles/prof_deepz.py <regime> <cells> [reps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import clr_b200  # noqa: E402,F401
from clr_b200 import runtime, workloads  # noqa: E402

regime = sys.argv[1] if len(sys.argv) > 1 else "stable"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = runtime.Context(0)
net = workloads.synthetic_relu_net()
grid = workloads.c4_grid(regime)
dn = ctx.upload(net)
dev = torch.device("cuda", 0)
cc = torch.from_numpy(np.concatenate(grid.centers)).to(dev)
rr = torch.from_numpy(np.concatenate(grid.radii)).to(dev)
lo = torch.empty((cells, 1), dtype=torch.float64, device=dev)
hi = torch.empty((cells, 1), dtype=torch.float64, device=dev)
pp = runtime.PrePostArg(None, None)
ctx.set_profiling(True)
for _ in range(reps):
    ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, 400000, cells, pp, lo, hi)
    ctx.synchronize()
    print(ctx.last_pass_info())
ctx.close()
