"""A/B timing of library builds on the C4 workload (10^6 cells, device-resident I/O, no statistics).

    python profiles/ab_bench.py <regimes, comma separated> <lib name>[,<lib name>...] [steps]

Every library (closedloopreachability.jl_b200/libclr_b200_<name>.so, or `main` for the product build) runs in a fresh
process; prints the per-pass device time of the DeepZ call (CUDA events of clr_set_profiling, L2 flushed between steps)
and a checksum of the boxes, so that variants can be told apart from broken builds.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "closedloopreachability.jl_b200")


def child(regime, steps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import clr_b200  # noqa: F401
    from clr_b200 import runtime, workloads
    ctx = runtime.Context(0)
    dev = torch.device("cuda", 0)
    net = workloads.synthetic_relu_net()
    g = workloads.c4_grid(regime)
    dn = ctx.upload(net)
    cnt = len(g)
    cc = torch.from_numpy(np.concatenate(g.centers)).to(dev)
    rr = torch.from_numpy(np.concatenate(g.radii)).to(dev)
    lo = torch.empty((cnt, 1), dtype=torch.float64, device=dev)
    hi = torch.empty_like(lo)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pp = runtime.PrePostArg(None, None)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.set_profiling(True)
    ms = np.zeros(3)
    for s in range(steps + 2):
        with torch.cuda.stream(stream):
            flush.zero_()
        ctx.deepz_boxgrid_dev(dn, 6, g.partition, cc, rr, 0, cnt, pp, lo, hi)
        ctx.synchronize()
        if s >= 2:
            ms += np.array(ctx.last_pass_info()["pass_ms"])
    ms /= steps
    info = ctx.last_pass_info()
    print(json.dumps({"regime": regime, "pass_ms": [round(float(x), 3) for x in ms], "total_ms": round(float(ms.sum()), 3),
                      "Mprop_s": round(cnt / ms.sum() / 1e3, 2), "retried": info["retried"], "big": info["big"],
                      "sum_lo": float(lo.sum().item()), "sum_hi": float(hi.sum().item())}))
    ctx.close()


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    regimes = sys.argv[1].split(",")
    libs = sys.argv[2].split(",")
    steps = sys.argv[3] if len(sys.argv) > 3 else "5"
    for name in libs:
        path = os.path.join(PKG, "libclr_b200.so" if name == "main" else f"libclr_b200_{name}.so")
        for r in regimes:
            env = dict(os.environ, CLR_B200_LIB=path)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", r, steps], env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else "ERR " + out.stderr[-400:]
            print(f"{name:12s} {line}", flush=True)
