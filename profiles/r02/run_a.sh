set -x
python -m pytest tests/test_gpu_deepz.py -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; tail -3 gpurun_out/r2a_tests.log
for r in stable mixed dense; do python bench.py --steps 10 --warmup 3 --regime $r --no-secondary --no-cpu-baseline > gpurun_out/r2a_bench_$r.json 2> gpurun_out/r2a_bench_$r.err; done
python - <<'PY'
import json
for r in ("stable","mixed","dense"):
    try:
        d=json.load(open(f"gpurun_out/r2a_bench_{r}.json"))
        print(r, d["value"], d["ms_per_step"], d["roofline"]["pass_ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(r, "ERR", e)
PY
