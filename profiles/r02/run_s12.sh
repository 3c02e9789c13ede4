python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python profiles/rollout_small.py 2>&1 | grep 'split=1' | tail -4
python - <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import clr_b200
from clr_b200 import runtime, workloads
ctx = runtime.Context(0)
net = workloads.synthetic_relu_net(); g = workloads.c4_grid("stable"); dn = ctx.upload(net)
for tag in ("pageable", "registered"):
    r = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, stats=False)
    t0 = time.perf_counter()
    for _ in range(5): r = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, stats=False)
    print(tag, "deepz_boxgrid 1e6 cells, fresh NumPy outputs:", round((time.perf_counter() - t0) / 5 * 1e3, 2), "ms per call")
    break
PY
