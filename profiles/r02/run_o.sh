CLR_TEST_DEVICES=0,1 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/r2o_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o_bench_n2.json"))
print("value", d["value"], "ms", d["ms_per_step"], "per_rank", d["per_rank"]["ms_per_step"], d["per_rank"]["cells"], "frac", d["roofline"]["frac"])
print("regimes", {k:(v["value"], v["rank_ms_per_step"]) for k,v in d.get("regimes",{}).items()})
print(d["config"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary --no-regimes --shards equal 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('equal shards:', d['value'], d['per_rank']['ms_per_step'])"
python - <<'PY'
# the library's own multi-GPU path on two real devices, timed
import sys, time; sys.path.insert(0, ".")
import numpy as np, clr_b200
from clr_b200 import runtime, workloads
m = runtime.MultiContext([0, 1]); net = workloads.synthetic_relu_net(); mn = m.upload(net)
from clr_b200.splitters import split_box
g = split_box([0.29] * 6, [0.31] * 6, [10] * 5 + [20])
for bal in (False, "cyclic"):
    for _ in range(3):
        t0 = time.perf_counter(); r = m.deepz_boxgrid(mn, g.partition, g.centers, g.radii, hull=True, stats=False, balance=bal); dt = time.perf_counter() - t0
    print("clr_multi_deepz_boxgrid 2e6 cells, balance", bal, f"{dt*1e3:.1f} ms wall = {2e6/dt/1e6:.1f} M/s incl. D2H of 32 MB; shard ms", r["shard_ms"], "hull", r["hull_lo"], r["hull_hi"])
PY
