python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python profiles/latency_small.py 2>&1 | tail -4
python bench.py --steps 5 --no-secondary > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("regimes", {k:(v["value"], v["ms_per_step"], v["roofline"]["frac"]) for k,v in d.get("regimes",{}).items()})
PY
