"""Regenerates the committed fixtures under tests/golden/ (run in the build container only).

* nets/*.npz   -- the bundled POLAR controllers named by BASELINE.json's configs, converted with
                  clr_b200.read_POLAR from /root/reference/models/*/*.polar (data, not source);
                  /root/reference does not exist on the GPU box, so tests read these instead.
* *_oracle.npz -- outputs of the NumPy oracle (oracle/deepz_np.py) on fixed inputs: regression
                  pins of the restatement (the reference itself publishes no numeric vectors,
                  SURVEY.md section 4 -- "parity unpinned").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import clr_b200  # noqa: E402
from oracle import deepz_np as dz  # noqa: E402

REF = "/root/reference/models"
HERE = os.path.dirname(os.path.abspath(__file__))
NETS = {
    "TORA_ReLU": "TORA/TORA_ReLU_controller.polar",
    "TORA_ReLUtanh": "TORA/TORA_ReLUtanh_controller.polar",
    "TORA_sigmoid": "TORA/TORA_sigmoid_controller.polar",
    "ACC_relu": "ACC/ACC_controller_relu.polar",
    "ACC_tanh": "ACC/ACC_controller_tanh.polar",
    "Unicycle": "Unicycle/Unicycle_controller.polar",
    "Quadrotor": "Quadrotor/Quadrotor_controller.polar",
    "InvertedPendulum": "InvertedPendulum/InvertedPendulum_controller.polar",
    "AttitudeControl": "AttitudeControl/AttitudeControl_controller.polar",
}
for i in range(1, 10):
    NETS[f"VerticalCAS_{i}"] = f"VerticalCAS/VerticalCAS_controller_{i}.polar"


def main():
    os.makedirs(os.path.join(HERE, "nets"), exist_ok=True)
    for name, rel in NETS.items():
        net = clr_b200.read_POLAR(os.path.join(REF, rel))
        net.save_npz(os.path.join(HERE, "nets", name + ".npz"))
        print(name, net.dims, [l.activation for l in net.layers])

    # TORA, X0 of models/TORA/TORA.jl:98, BoxSplitter([4,4,3,5]) of :188 and the 10^3 split of C2
    tora = clr_b200.FeedforwardNetwork.load_npz(os.path.join(HERE, "nets", "TORA_ReLU.npz")).as_tuples()
    lo, hi = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]
    out = {}
    for tag, part in (("p4435", [4, 4, 3, 5]), ("p1010101", [10, 10, 10, 1])):
        st = {}
        l, h = dz.deepz_boxgrid(tora, lo, hi, part, post=("shift", -10.0), stats=st)
        out[f"{tag}_lo"], out[f"{tag}_hi"] = l, h
        out[f"{tag}_p_in"] = np.array([s["p_in"] for s in st["cells"]])
    c, G = dz.hyperrectangle_to_zonotope((np.array(lo) + hi) / 2, (np.array(hi) - lo) / 2)
    _, _, l, h = dz.controller_forward(tora, c, G, None, ("shift", -10.0))
    out["whole_lo"], out["whole_hi"] = l, h
    np.savez_compressed(os.path.join(HERE, "tora_oracle.npz"), **out)

    # Quadrotor (sigmoid), X0 of models/Quadrotor/Quadrotor.jl:98-99, 2^6 cells, full output zonotope of cell 0
    quad = clr_b200.FeedforwardNetwork.load_npz(os.path.join(HERE, "nets", "Quadrotor.npz")).as_tuples()
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    l, h = dz.deepz_boxgrid(quad, -r, r, [2] * 6 + [1] * 6)
    cells, rad = dz.split_box(-r, r, [2] * 6 + [1] * 6)
    c, G = dz.hyperrectangle_to_zonotope(cells[0], rad)
    c, G = dz.deepz_forward(quad, c, G)
    np.savez_compressed(os.path.join(HERE, "quadrotor_oracle.npz"), lo=l, hi=h, cell0_c=c, cell0_G=G)

    # order-2 / order-3 reduction of that output zonotope (src/setops.jl:84-86) and _project_oa
    # (src/setops.jl:9-12) of a fixed batch of Taylor models (seeded; TMJets-like orders 4 / 2 in 3 variables)
    out = {}
    for order in (1, 2, 3):
        rc, rG = dz.reduce_order_gir05(c, G, order)
        out[f"red{order}_c"], out[f"red{order}_G"] = rc, rG
    rng = np.random.default_rng(77)
    n, oT, oQ, N = 3, 4, 2, 6
    exps = dz.monomial_table(n, oQ)
    deg = exps.sum(axis=1)
    coeffs = rng.normal(size=(N, n, oT + 1, exps.shape[0])) * (0.1 ** deg)[None, None, None, :] / (1.0 + np.arange(oT + 1))[None, None, :, None]
    rem = np.sort(rng.normal(size=(N, n, 2)) * 1e-6, axis=2)
    dt = rng.uniform(0.05, 0.2, size=N)
    out.update(tm_exps=exps, tm_coeffs=coeffs, tm_rem=rem, tm_dt=dt)
    for i in range(N):
        pc, pG = dz.project_oa_taylor_model(exps, coeffs[i], rem[i], dt[i], [0, 2])
        out[f"proj{i}_c"], out[f"proj{i}_G"] = pc, pG
    np.savez_compressed(os.path.join(HERE, "setops_oracle.npz"), **out)


if __name__ == "__main__":
    main()
