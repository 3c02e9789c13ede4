python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "peak", d["roofline"]["peak"])
print("regimes", {k:(v["value"], v["ms_per_step"], v["roofline"]["frac"]) for k,v in d.get("regimes",{}).items()})
print("cpu", d.get("cpu_baseline")); print("secondary", {k:v.get("value") for k,v in d["secondary"].items() if isinstance(v,dict)})
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
for v in main r3b5 r2b6 r2b8; do echo "== rollout $v"; CLR_B200_LIB=$PWD/closedloopreachability.jl_b200/libclr_b200$( [ $v = main ] || echo _$v ).so python profiles/rollout_bench.py 1000000 2>&1 | tail -3; done | tee gpurun_out/r2l_rollout.txt
