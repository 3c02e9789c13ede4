"""Copy the files written by profiles/r02_final_run.sh (gpurun_out/f2_*) into profiles/r02/ and print the headline numbers.

    python profiles/collect_r02.py [round dir, default r02]
"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02")
os.makedirs(P, exist_ok=True)
pairs = {"f2_bench.json": "bench_n1.json", "f2_bench_ref.json": "bench_reference_n1.json",
         "f2_latency.txt": "latency_small_b200.txt", "f2_reduce.txt": "reduce_bench_b200.txt",
         "f2_project.txt": "project_bench_b200.txt", "f2_shards.txt": "shard_costs_b200.txt",
         "f2_ncu_quad.json": "ncu_deepz_c3_quadrotor.json", "f2_ncu_quad_lines.txt": "ncu_deepz_c3_quadrotor_lines.txt",
         "f2_ncu_controller_cw.json": "ncu_controller_cw.json", "f2_ncu_controller_cw_lines.txt": "ncu_controller_cw_lines.txt",
         "f2_ncu_rollout_ode.json": "ncu_rollout_ode.json", "f2_ncu_rollout_ode_lines.txt": "ncu_rollout_ode_lines.txt",
         "f2_latency_c.txt": "latency_c_abi_b200.txt", "f2_plain_rollout.log": "rollout_bench_b200.txt",
         "f2_ncu_reduce.json": "ncu_reduce_reg_launch.json", "f2_ncu_reduce_lines.txt": "ncu_reduce_reg_launch_lines.txt",
         "f2_ncu_wide.json": "ncu_deepz_wide_pass.json", "f2_ncu_wide_lines.txt": "ncu_deepz_wide_pass_lines.txt",
         "f2_pytest.log": "pytest_gpu.log", "f2_bench_n2.json": "bench_n2.json", "f2_bench_n4.json": "bench_n4.json",
         "f2_bench_n8.json": "bench_n8.json"}
for r in ("stable", "mixed", "dense"):
    pairs[f"f2_ncu_deepz_{r}.json"] = f"ncu_deepz_{r}.json"
    pairs[f"f2_ncu_deepz_{r}_lines.txt"] = f"ncu_deepz_{r}_lines.txt"
for src, dst in pairs.items():
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copy(s, os.path.join(P, dst))
    else:
        print("missing:", src)

lc = os.path.join(G, "f2_launches.csv")
if os.path.exists(lc):
    with open(lc) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    with open(os.path.join(P, "ncu_launches_bench.csv"), "w") as fh:
        fh.writelines(lines)
    rows = list(csv.reader(lines))
    ci = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(rows[0]) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[ci["Metric Unit"]], 1)
        a = agg.setdefault(r[ci["Kernel Name"]][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "ncu_launches_bench_summary.csv"), "w") as fh:
        fh.write("kernel,launches,total_ms,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f'"{k}",{a[0]},{a[1]:.3f},{a[1] / tot:.3f}\n')

d = json.load(open(os.path.join(G, "f2_bench.json")))


def show(name, v, ms, e2e, rf):
    print(f"{name:7s} {v / 1e6:6.2f} M/s  {ms:7.2f} ms  e2e {e2e / 1e6:6.2f}  frac {rf['frac']:.3f}  {rf['achieved']:.2f} TF  "
          f"passes {[round(x, 2) for x in rf['pass_ms_per_step']]}  retried {rf['cells_retried']}  big {rf['cells_big']}")


show(d["config"]["regime"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"])
for k, v in d.get("regimes", {}).items():
    show(k, v["value"], v["ms_per_step"], v["e2e"], v["roofline"])
print("peak", d["roofline"]["peak"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "clocks", d["clocks"],
      "launches", d["gpu_launches"])
print({k: (round(v["value"] / 1e6, 2), round(v["ms"], 2), round(v["frac_of_fp64_fma_peak"], 3))
       for k, v in d["secondary"].items() if isinstance(v, dict) and "ms" in v})
for r in ("stable", "mixed", "dense"):
    f = os.path.join(G, f"f2_ncu_deepz_{r}.json")
    if os.path.exists(f):
        j = json.load(open(f))
        print(r, j["kernel"][5:32], "dmma", round(j["dmma_pipe_active_pct"], 1), "issue", round(j["issue_active_pct"], 1),
              "traffic", j["traffic"], "ms", round(j["gpu_time_ms"], 2))
