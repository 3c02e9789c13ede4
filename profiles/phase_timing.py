"""Per-phase cycle breakdown of deepz_kernel (needs the -DDZ_TIMING build: CLR_B200_LIB=.../libclr_b200_timing.so)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clr_b200
from clr_b200 import runtime, workloads
regime = sys.argv[1] if len(sys.argv) > 1 else "stable"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
ctx = runtime.Context(0)
net = workloads.synthetic_relu_net(); grid = workloads.c4_grid(regime); dn = ctx.upload(net)
dev = torch.device("cuda", 0)
cc = torch.from_numpy(np.concatenate(grid.centers)).to(dev); rr = torch.from_numpy(np.concatenate(grid.radii)).to(dev)
lo = torch.empty((cells, 1), dtype=torch.float64, device=dev); hi = torch.empty_like(lo)
pp = runtime.PrePostArg(None, None)
ctx.set_profiling(True)
names = ["input", "product", "appended+sync1", "relax+sync2", "layout+sync3", "output", "fetch", "layout-scan"]
for T in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"])]:
    ctx.set_tile_cells(T)
    for _ in range(2):
        ctx.deepz_boxgrid_dev(dn, 6, grid.partition, cc, rr, 400000, cells, pp, lo, hi)
        ctx.synchronize()
    t = (C.c_ulonglong * 8)()
    ctx.L.clr_debug_timers.argtypes = [C.c_void_p, C.c_void_p]
    ctx.L.clr_debug_timers(ctx.h, t)
    tot = sum(t)
    print(f"T={T}", ctx.last_pass_info()["pass_ms"][0], "ms;", ", ".join(f"{n} {100*v/tot:.1f}%" for n, v in zip(names, t)))
ctx.close()
