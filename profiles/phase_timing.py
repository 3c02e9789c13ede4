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
if regime == "quadrotor":      # C3: 12 -> 64^3 -> 3 sigmoid controller, 4^6 cells of X0 (models/Quadrotor/Quadrotor.jl:98-99)
    net = clr_b200.FeedforwardNetwork.load_npz(os.path.join(ROOT, "tests", "golden", "nets", "Quadrotor.npz"))
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    grid = clr_b200.split_box(-r, r, [4] * 6 + [1] * 6)
    begin, n_in, m_out = 0, 12, 3
    cells = min(cells, 4096)
else:
    net = workloads.synthetic_relu_net(); grid = workloads.c4_grid(regime)
    begin, n_in, m_out = 400000, 6, 1
dn = ctx.upload(net)
dev = torch.device("cuda", 0)
cc = torch.from_numpy(np.concatenate(grid.centers)).to(dev); rr = torch.from_numpy(np.concatenate(grid.radii)).to(dev)
lo = torch.empty((cells, m_out), dtype=torch.float64, device=dev); hi = torch.empty_like(lo)
pp = runtime.PrePostArg(None, None)
ctx.set_profiling(True)
names = ["input", "product", "appended+sync1", "relax+sync2", "layout+sync3", "output", "fetch", "layout-scan"]
for T in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"])]:
    ctx.set_tile_cells(T)
    for _ in range(2):
        ctx.deepz_boxgrid_dev(dn, n_in, grid.partition, cc, rr, begin, cells, pp, lo, hi)
        ctx.synchronize()
    t = (C.c_ulonglong * 8)()
    ctx.L.clr_debug_timers.argtypes = [C.c_void_p, C.c_void_p]
    ctx.L.clr_debug_timers(ctx.h, t)
    tot = sum(t)
    print(f"T={T}", ctx.last_pass_info()["pass_ms"][0], "ms;", ", ".join(f"{n} {100*v/tot:.1f}%" for n, v in zip(names, t)))
ctx.close()
