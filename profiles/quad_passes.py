import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import clr_b200
from clr_b200 import runtime
ctx = runtime.Context(0)
net = clr_b200.FeedforwardNetwork.load_npz("/root/repo/tests/golden/nets/Quadrotor.npz")
dq = ctx.upload(net)
r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
g = clr_b200.split_box(-r, r, [4] * 6 + [1] * 6)
ctx.set_profiling(True)
for T in (0, 1, 2, 3):
    ctx.set_tile_cells(T)
    for _ in range(3):
        t0 = time.perf_counter()
        out = ctx.deepz_boxgrid(dq, g.partition, g.centers, g.radii, stats=False)
        dt = time.perf_counter() - t0
    print(T, f"{dt*1e6:.0f} us", ctx.last_pass_info())
