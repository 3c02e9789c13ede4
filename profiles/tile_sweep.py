"""Tile-size sweep of the DeepZ main pass on a slice of the C4 workload (device-resident I/O)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import runtime, workloads
regime = sys.argv[1] if len(sys.argv) > 1 else "stable"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
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
for T in [0, 1, 2, 3, 4, 6, 8, 10, 12, 13, 14, 16, 20, 24, 32]:
    ctx.set_tile_cells(T)
    for _ in range(2):
        ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, 400000, cells, pp, lo, hi)
        ctx.synchronize()
    print(T, ctx.last_pass_info(), flush=True)
ctx.close()
