"""The bench.py contract the driver relies on: ONE JSON line on stdout with the agreed keys.  The reference arm (the C
restatement timed on the host cores) runs without a GPU, so its line is checked here; the GPU arm is checked on a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _one_line(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]          # library banners must not reach stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _one_line(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "propagations/s" and d["value"] > 0 and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    d = _one_line(["--steps", "2", "--warmup", "3", "--no-secondary", "--no-cpu-baseline", "--cells-per-dim", "8"])
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["clocks"]["sm_mhz"] is None or d["clocks"]["sm_mhz"] > 0
    # the whole-step roofline can be recomputed from the line itself, the other regimes and the per-rank
    # times are on the same line, and the FP64 peak was measured in this run
    assert abs(r["frac"] - r["algorithmic_flops_per_step"] / (d["ms_per_step"] * 1e-3) / 1e12 / r["peak"]) < 1e-6 * r["frac"] + 1e-12
    assert 20.0 < r["peak"] < 60.0 and "measured in this run" in r["peak_source"]
    assert set(d["regimes"]) == {"mixed", "dense"} and all(v["value"] > 0 and 0 < v["roofline"]["frac"] < 1 for v in d["regimes"].values())
    assert d["per_rank"]["ms_per_step"] == [d["ms_per_step"]] and d["per_rank"]["cells"] == [8 ** 6]
