import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import clr_b200
from clr_b200 import runtime, workloads
ctx = runtime.Context(0)
net = clr_b200.FeedforwardNetwork.load_npz("/root/repo/tests/golden/nets/TORA_ReLU.npz")
dt = ctx.upload(net)
g = clr_b200.split_box(*workloads.TORA_X0, [10, 10, 10, 1])
ctx.set_profiling(True)
for T in (0, 4, 7):
    ctx.set_tile_cells(T)
    for _ in range(5):
        t0 = time.perf_counter()
        out = ctx.deepz_boxgrid(dt, g.partition, g.centers, g.radii, post=("shift", -10.0), stats=False)
        d = time.perf_counter() - t0
    print(T, f"{d*1e6:.0f} us", ctx.last_pass_info())
