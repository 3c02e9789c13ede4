"""Copy the files written by profiles/r01_final_run.sh (gpurun_out/f_*) into profiles/<round>/ and print the headline numbers."""
import collections, csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r01")
os.makedirs(P, exist_ok=True)
pairs = {"f_bench_stable.json": ["bench_stable_final.json", "bench_default_n1.json"], "f_bench_mixed.json": ["bench_mixed_final.json"],
         "f_bench_dense.json": ["bench_dense_final.json"], "f_bench_ref.json": ["bench_reference_n1.json"],
         "f_ncu_deepz.json": ["ncu_deepz_bench_launch.json"], "f_ncu_deepz_lines.txt": ["ncu_deepz_bench_launch_lines.txt"],
         "f_ncu_rollout.json": ["ncu_rollout_launch.json"], "f_ncu_rollout_lines.txt": ["ncu_rollout_launch_lines.txt"],
         "f_ncu_reduce.json": ["ncu_reduce_launch.json"], "f_ncu_reduce_lines.txt": ["ncu_reduce_launch_lines.txt"]}
for src, dsts in pairs.items():
    for d in dsts:
        shutil.copy(os.path.join(G, src), os.path.join(P, d))
with open(os.path.join(G, "f_launches.csv")) as fh:
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
    a = agg.setdefault(r[ci["Kernel Name"]][:90], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "ncu_launches_bench_summary.csv"), "w") as fh:
    fh.write("kernel,launches,total_ms,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'"{k}",{a[0]},{a[1]:.3f},{a[1] / tot:.3f}\n')
for reg in ("stable", "mixed", "dense"):
    d = json.load(open(os.path.join(G, f"f_bench_{reg}.json"))); r = d["roofline"]
    print(reg, f"{d['value'] / 1e6:.2f} M/s  e2e {d['e2e']['value'] / 1e6:.2f}  frac {r['frac']:.3f}  {r['achieved']:.2f} TF  kernel {r['kernel_ms_per_step']:.2f} ms"
          f"  cpu {d['cpu_baseline']['value'] / 1e3:.1f} k/s  clocks {d['clocks']['sm_mhz']}  launches {d['gpu_launches']}")
s = json.load(open(os.path.join(G, "f_bench_stable.json")))["secondary"]
print({k: (round(v["value"] / 1e6, 2), round(v["ms"], 2), round(v["frac_of_fp64_fma_peak"], 3)) for k, v in s.items() if isinstance(v, dict)})
print(open(os.path.join(G, "f_ncu_deepz.json")).read())
