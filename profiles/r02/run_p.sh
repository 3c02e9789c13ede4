CLR_TEST_DEVICES=0,1 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
# the library's own multi-GPU path on two real devices, timed
import sys, time; sys.path.insert(0, ".")
import numpy as np, clr_b200
from clr_b200 import runtime, workloads
m = runtime.MultiContext([0, 1]); net = workloads.synthetic_relu_net(); mn = m.upload(net)
from clr_b200.splitters import split_box
g = split_box([0.29] * 6, [0.31] * 6, [10] * 5 + [20])
for bal in (False, "cost", "cyclic"):
    for _ in range(3):
        t0 = time.perf_counter(); r = m.deepz_boxgrid(mn, g.partition, g.centers, g.radii, hull=True, stats=False, balance=bal); dt = time.perf_counter() - t0
    print("clr_multi_deepz_boxgrid 2e6 cells, balance", bal, f"{dt*1e3:.1f} ms wall = {2e6/dt/1e6:.1f} M/s incl. D2H of 32 MB into pageable memory; shard ms", r["shard_ms"], "hull", r["hull_lo"], r["hull_hi"])
PY
