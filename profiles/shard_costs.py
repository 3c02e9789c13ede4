"""Time and algorithmic FLOPs of each rank's shard of the N-GPU C4 workload, one after the other on ONE GPU
(calibrates the cost model behind parallel.balanced_bounds).   python profiles/shard_costs.py <world> [regime]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import _lib, parallel, runtime, workloads
from clr_b200.splitters import split_box
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
regime = sys.argv[2] if len(sys.argv) > 2 else "stable"
c0, s = {"dense": (0.0, 1.0), "mixed": (0.3, 0.1), "stable": (0.3, 0.01)}[regime]
grid = split_box([c0 - s] * 6, [c0 + s] * 5 + [c0 - s + 2 * s * world], [10] * 5 + [10 * world])
net = workloads.synthetic_relu_net()
ctx = runtime.Context(0)
dn = ctx.upload(net)
dev = torch.device("cuda", 0)
cc = torch.from_numpy(np.concatenate(grid.centers)).to(dev); rr = torch.from_numpy(np.concatenate(grid.radii)).to(dev)
pp = runtime.PrePostArg(None, None)
ctx.set_profiling(True)
def run(b, cnt):
    lo = torch.empty((cnt, 1), dtype=torch.float64, device=dev); hi = torch.empty_like(lo)
    st = _lib.Stats()
    ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, b, cnt, pp, lo, hi, stats=st); ctx.synchronize()
    ms = []
    for _ in range(4):
        ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, b, cnt, pp, lo, hi); ctx.synchronize()
        ms.append(sum(ctx.last_pass_info()["pass_ms"]))
    return float(np.median(ms[1:])), st.flops
print("equal-count shards")
for r in range(world):
    b, cnt = parallel.shard_range(len(grid), r, world)
    ms, fl = run(b, cnt)
    print(f"rank {r}: cells {cnt} ms {ms:.3f} flops/cell {fl / cnt:.0f} ns/cell {1e6 * ms / cnt:.3f}")
for base in (0.0, 2.0e5, 5.0e5):
    bounds, cost = parallel.balanced_bounds(ctx, dn, grid, world, base_cost=base)
    print(f"cost-weighted shards, base_cost {base:g}: bounds", bounds)
    for r in range(world):
        ms, fl = run(bounds[r], bounds[r + 1] - bounds[r])
        print(f"rank {r}: cells {bounds[r + 1] - bounds[r]} ms {ms:.3f}")
ctx.close()
