"""GPU parity tests of the simulate() path (clr_rollout / clr_forward_points) against the CPU oracle.
Tolerance (north_star): trajectories within 1e-10 for identical inputs; step counts identical."""
import numpy as np
import pytest

from conftest import load_net

pytestmark = pytest.mark.gpu
TRAJ_ATOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    from clr_b200 import runtime
    c = runtime.Context(0)
    yield c
    c.close()


def _oracle():
    from oracle import cbind
    return cbind


@pytest.mark.parametrize("name,pre,post", [("Unicycle", None, ("shift", -20.0)), ("TORA_ReLU", None, ("shift", -10.0)),
                                            ("TORA_sigmoid", None, ("scale", 11.0)), ("Quadrotor", None, None),
                                            ("AttitudeControl", None, ("project", [2, 0]))])
def test_forward_points(ctx, name, pre, post):
    net = load_net(name)
    X = np.random.default_rng(3).uniform(-1, 1, (1000, net.dim_in))
    got = ctx.forward_points(ctx.upload(net), X, pre=pre, post=post)
    ref = _oracle().forward_points(_oracle().Net(net.as_tuples()), X, pre=pre, post=post)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("alg", ["Tsit5", "RK4"])
def test_unicycle_rollout(ctx, alg):
    from clr_b200 import workloads
    net = load_net("Unicycle")
    rng = np.random.default_rng(5)
    N, iters = 2000, 50
    x0, Wd = workloads.unicycle_samples(rng, N, iters)
    got = ctx.rollout(ctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 10.0, iters, alg=alg, post=("shift", -20.0))
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), "Unicycle", 4, 1, 2, x0, Wd, 0.0, 0.2, 10.0, iters, alg=alg,
                            post=("shift", -20.0))
    assert np.array_equal(got["naccept"], ref["naccept"]) and np.array_equal(got["nreject"], ref["nreject"])
    np.testing.assert_allclose(got["X_end"], ref["X_end"], rtol=0, atol=TRAJ_ATOL)
    np.testing.assert_allclose(got["U"], ref["U"], rtol=0, atol=TRAJ_ATOL)


def test_step_log_matches_oracle(ctx):
    from clr_b200 import workloads
    net = load_net("Unicycle")
    x0, Wd = workloads.unicycle_samples(np.random.default_rng(6), 64, 4)
    got = ctx.rollout(ctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 0.75, 4, post=("shift", -20.0), log_cap=16)
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), "Unicycle", 4, 1, 2, x0, Wd, 0.0, 0.2, 0.75, 4,
                            post=("shift", -20.0), log_cap=16)
    assert np.array_equal(got["log_n"], ref["log_n"])
    for it in range(4):
        for j in range(64):
            k = ref["log_n"][it, j]
            np.testing.assert_allclose(got["log_t"][it, j, :k], ref["log_t"][it, j, :k], rtol=0, atol=1e-13)
            np.testing.assert_allclose(got["log_u"][it, j, :k], ref["log_u"][it, j, :k], rtol=0, atol=TRAJ_ATOL)
    # the last period is shorter: it ends at T = 0.75 (src/simulate.jl:126)
    assert abs(got["log_t"][3, 0, got["log_n"][3, 0] - 1] - 0.75) < 1e-15


@pytest.mark.parametrize("plant,name,post,lo,hi", [
    ("TORA", "TORA_ReLU", ("shift", -10.0), [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]),
    ("InvertedPendulum", "InvertedPendulum", None, [1.0, 0.0], [1.2, 0.2])])
def test_other_plants(ctx, plant, name, post, lo, hi):
    net = load_net(name)
    rng = np.random.default_rng(8)
    x0 = rng.uniform(lo, hi, (500, len(lo)))
    tau = 1.0 if plant == "TORA" else 0.05
    got = ctx.rollout(ctx.upload(net), plant, x0, None, 0.0, tau, 10 * tau, 10, post=post)
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), plant, len(lo), 0, 1, x0, None, 0.0, tau, 10 * tau, 10, post=post)
    assert np.array_equal(got["naccept"], ref["naccept"])
    np.testing.assert_allclose(got["X_end"], ref["X_end"], rtol=0, atol=TRAJ_ATOL)


def test_rollout_argument_errors(ctx):
    from clr_b200 import runtime
    net = load_net("Unicycle")
    dn = ctx.upload(net)
    x0 = np.zeros((4, 4))
    with pytest.raises(runtime.ClrArgumentError):
        ctx.rollout(dn, "TORA", x0, None, 0.0, 0.2, 1.0, 5)                # 2 controller outputs, TORA takes 1
    with pytest.raises(runtime.ClrError):
        ctx.rollout(dn, "Unicycle", x0, np.zeros((5, 4, 1)), 0.0, 0.2, 1.0, 5, post=("shift", -20.0), log_cap=1)


def test_unicycle_c5_full_size(ctx):
    """BASELINE config C5: 10^6 trajectories x 50 periods.  A sample strided over the whole batch is checked against the oracle; the whole
    batch through shard consistency (a shard run alone reproduces its slice bit for bit) and finiteness."""
    from clr_b200 import workloads
    net = load_net("Unicycle")
    dn = ctx.upload(net)
    N, iters = 1000000, 50
    x0, Wd = workloads.unicycle_samples(np.random.default_rng(5), N, iters)
    got = ctx.rollout(dn, "Unicycle", x0, Wd, 0.0, 0.2, 10.0, iters, post=("shift", -20.0), counts=False)
    assert np.all(np.isfinite(got["X_end"]))
    # parity sample strided over ALL 10^6 trajectories (every block of the grid, every position inside a thread's group of
    # four), plus a contiguous run at each end
    idx = np.unique(np.concatenate([np.arange(0, N, 331), np.arange(600), np.arange(N - 600, N)]))
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), "Unicycle", 4, 1, 2, x0[idx], Wd[:, idx], 0.0, 0.2, 10.0, iters,
                            post=("shift", -20.0))
    assert len(idx) > 4000
    np.testing.assert_allclose(got["X_end"][:, idx], ref["X_end"], rtol=0, atol=TRAJ_ATOL)
    np.testing.assert_allclose(got["U"][:, idx], ref["U"], rtol=0, atol=TRAJ_ATOL)
    b, e = 375000, 500000                                  # rank 3 of 8
    part = ctx.rollout(dn, "Unicycle", x0[b:e], Wd[:, b:e], 0.0, 0.2, 10.0, iters, post=("shift", -20.0), counts=False)
    assert np.array_equal(part["X_end"], got["X_end"][:, b:e]) and np.array_equal(part["U"], got["U"][:, b:e])


def test_simulate_mirror(ctx):
    """simulate(prob; T, trajectories, include_vertices) (src/simulate.jl:68-143) on the Unicycle model
    (models/Unicycle/Unicycle.jl:34-80) against the oracle fed with the same samples."""
    import clr_b200
    net = load_net("Unicycle")
    vars_ = {"states": [1, 2, 3, 4], "disturbances": [5], "controls": [6, 7]}
    lo = [9.5, -4.5, 2.1, 1.5, -1e-4, 0.0, 0.0]
    hi = [9.55, -4.45, 2.11, 1.51, 1e-4, 0.0, 0.0]
    prob = clr_b200.ControlledPlant("Unicycle", lo, hi, net, vars_, 0.2,
                                    postprocessing=clr_b200.UniformAdditivePostprocessing(-20.0))
    sol = clr_b200.simulate(ctx, prob, T=1.5, trajectories=40, include_vertices=True, rng=np.random.default_rng(3),
                            save_steps=True)
    assert len(sol) == 40 + 16                       # the vertices of X0 are appended (src/simulate.jl:89-90)
    iters = 8                                        # ceil(1.5 / 0.2); the last period is shorter
    assert sol.X_end.shape == (iters, 56, 4) and sol.controls().shape == (56, iters, 2)
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), "Unicycle", 4, 1, 2, sol.x0, sol.Wd, 0.0, 0.2, 1.5, iters,
                            post=("shift", -20.0))
    np.testing.assert_allclose(sol.X_end, ref["X_end"], rtol=0, atol=TRAJ_ATOL)
    one = sol[3]
    assert len(one) == iters and one.controls().shape == (iters, 2) and one.disturbances().shape == (iters, 1)
    t_last, u_last = one.trajectory()[-1]
    assert abs(t_last[-1] - 1.5) < 1e-15 and u_last.shape[1] == 7
    np.testing.assert_allclose(u_last[-1, :4], sol.X_end[-1, 3], rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        clr_b200.simulate(ctx, prob, trajectories=4)


UNICYCLE_RHS = """
    double sn, cs;
    sincos(x[2], &sn, &cs);
    dx[0] = x[3] * cs;
    dx[1] = x[3] * sn;
    dx[2] = q[2];
    dx[3] = q[1] + q[0];
"""
ATTITUDE_RHS = """
    const double xi = x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    dx[0] = 0.25 * (q[0] + x[1] * x[2]);
    dx[1] = 0.5 * (q[1] - 3 * x[0] * x[2]);
    dx[2] = q[2] + 2 * x[0] * x[1];
    dx[3] = 0.5 * (x[1] * (xi - x[5]) + x[2] * (xi + x[4]) + x[0] * (xi + 1));
    dx[4] = 0.5 * (x[0] * (xi + x[5]) + x[2] * (xi - x[3]) + x[1] * (xi + 1));
    dx[5] = 0.5 * (x[0] * (xi - x[4]) + x[1] * (xi + x[3]) + x[2] * (xi + 1));
"""


def test_runtime_compiled_plants(ctx):
    """clr_plant_compile / clr_rollout_plant: a user right-hand side compiled with NVRTC.  The Unicycle source must
    reproduce the built-in plant bit for bit; AttitudeControl (models/AttitudeControl/AttitudeControl.jl:34-57, a plant
    that is not built in) is checked against the oracle."""
    from clr_b200 import runtime, workloads
    uni = load_net("Unicycle")
    du = ctx.upload(uni)
    x0, Wd = workloads.unicycle_samples(np.random.default_rng(21), 3000, 10)
    plant = ctx.compile_plant(4, 1, 2, UNICYCLE_RHS)
    a = ctx.rollout(du, "Unicycle", x0, Wd, 0.0, 0.2, 2.0, 10, post=("shift", -20.0))
    b = ctx.rollout(du, plant, x0, Wd, 0.0, 0.2, 2.0, 10, post=("shift", -20.0))
    assert np.array_equal(a["X_end"], b["X_end"]) and np.array_equal(a["U"], b["U"]) and np.array_equal(a["naccept"], b["naccept"])
    att = load_net("AttitudeControl")
    da = ctx.upload(att)
    lo = np.array([-0.45, -0.55, 0.65, -0.75, 0.85, -0.65])          # AttitudeControl.jl:72-73
    hi = np.array([-0.44, -0.54, 0.66, -0.74, 0.86, -0.64])
    xa = np.random.default_rng(22).uniform(lo, hi, (1500, 6))
    pa = ctx.compile_plant(6, 0, 3, ATTITUDE_RHS)
    got = ctx.rollout(da, pa, xa, None, 0.0, 0.1, 3.0, 30)
    ref = _oracle().rollout(_oracle().Net(att.as_tuples()), "AttitudeControl", 6, 0, 3, xa, None, 0.0, 0.1, 3.0, 30)
    assert np.array_equal(got["naccept"], ref["naccept"])
    np.testing.assert_allclose(got["X_end"], ref["X_end"], rtol=0, atol=TRAJ_ATOL)
    np.testing.assert_allclose(got["U"], ref["U"], rtol=0, atol=TRAJ_ATOL)
    with pytest.raises(runtime.ClrArgumentError):
        ctx.rollout(da, plant, xa, None, 0.0, 0.1, 3.0, 30)          # 6 state columns into a 4-state plant


def _synthetic_net(dims, acts, seed):
    import clr_b200
    rng = np.random.default_rng(seed)
    layers = [(rng.normal(0, 1.0 / np.sqrt(n), (m, n)), rng.normal(0, 0.3, m), a)
              for n, m, a in zip(dims[:-1], dims[1:], acts)]
    return clr_b200.FeedforwardNetwork.from_tuples(layers)


# narrow controllers (every materialised vector has at most 8 entries) are evaluated from the constant bank by
# controller_cw_kernel: every record shape (4/2, 4/8, 8/8), 1-4 layers (pair, pair + pair, pair + single tail), every activation
NARROW = [([3, 40, 2], ["ReLU", "Id"]), ([4, 30, 4], ["Sigmoid", "Tanh"]), ([4, 25, 3, 7], ["ReLU", "ReLU", "Id"]),
          ([4, 30, 6, 20, 5], ["Tanh", "ReLU", "ReLU", "Sigmoid"]), ([8, 8], ["ReLU"]), ([2, 512, 1], ["ReLU", "Id"])]


@pytest.mark.parametrize("case", range(len(NARROW)))
def test_forward_points_narrow_networks(ctx, case):
    dims, acts = NARROW[case]
    net = _synthetic_net(dims, acts, 40 + case)
    for N in (1, 511, 1000):                      # not a multiple of the 512 points a block takes
        X = np.random.default_rng(50 + N).uniform(-1, 1, (N, dims[0]))
        got = ctx.forward_points(ctx.upload(net), X)
        ref = _oracle().forward_points(_oracle().Net(net.as_tuples()), X)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13)


def test_constant_bank_is_shared_between_contexts(ctx):
    """The packed controller of a split rollout lives in ONE constant bank per device: two contexts (two streams) that
    alternate between different controllers must each see their own."""
    from clr_b200 import runtime, workloads
    other = runtime.Context(0)
    try:
        uni = load_net("Unicycle")
        alt = _synthetic_net([4, 60, 2], ["ReLU", "Id"], 77)
        x0, Wd = workloads.unicycle_samples(np.random.default_rng(9), 700, 6)
        refs = [_oracle().rollout(_oracle().Net(n.as_tuples()), "Unicycle", 4, 1, 2, x0, Wd, 0.0, 0.2, 1.2, 6, post=p)
                for n, p in ((uni, ("shift", -20.0)), (alt, None))]
        du, da = ctx.upload(uni), other.upload(alt)
        for _ in range(3):
            a = ctx.rollout(du, "Unicycle", x0, Wd, 0.0, 0.2, 1.2, 6, post=("shift", -20.0))
            b = other.rollout(da, "Unicycle", x0, Wd, 0.0, 0.2, 1.2, 6)
            np.testing.assert_allclose(a["X_end"], refs[0]["X_end"], rtol=0, atol=TRAJ_ATOL)
            np.testing.assert_allclose(b["X_end"], refs[1]["X_end"], rtol=0, atol=TRAJ_ATOL)
            np.testing.assert_allclose(b["U"], refs[1]["U"], rtol=0, atol=TRAJ_ATOL)
    finally:
        other.close()


def test_split_rollout_equals_fused_kernel(ctx, monkeypatch):
    """One controller launch + one plant launch per period (the default for narrow controllers) against the fused
    rollout_blocked_kernel (CLR_B200_RO_SPLIT=0): the same arithmetic per trajectory, so the same bits."""
    from clr_b200 import runtime, workloads
    net = load_net("Unicycle")
    x0, Wd = workloads.unicycle_samples(np.random.default_rng(31), 5000, 12)
    monkeypatch.setenv("CLR_B200_RO_SPLIT", "0")
    fused_ctx = runtime.Context(0)
    monkeypatch.delenv("CLR_B200_RO_SPLIT")
    try:
        for alg in ("Tsit5", "RK4"):
            a = ctx.rollout(ctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 2.3, 12, alg=alg, post=("shift", -20.0))
            b = fused_ctx.rollout(fused_ctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 2.3, 12, alg=alg, post=("shift", -20.0))
            assert np.array_equal(a["X_end"], b["X_end"]) and np.array_equal(a["U"], b["U"])
            assert np.array_equal(a["naccept"], b["naccept"]) and np.array_equal(a["nreject"], b["nreject"])
    finally:
        fused_ctx.close()


def test_pageable_and_registered_host_buffers_agree(ctx):
    """Host buffers may be pageable (the driver stages the copy) or page-locked in place (clr_host_register: one DMA): the same
    bytes either way; registering twice / unregistering twice is harmless."""
    from clr_b200 import workloads
    net = load_net("Unicycle")
    dn = ctx.upload(net)
    N, iters = 70001, 9
    x0, Wd = workloads.unicycle_samples(np.random.default_rng(17), N, iters)
    a = ctx.rollout(dn, "Unicycle", x0, Wd, 0.0, 0.2, 1.8, iters, post=("shift", -20.0))
    ctx.host_register(x0); ctx.host_register(Wd); ctx.host_register(x0)
    try:
        b = ctx.rollout(dn, "Unicycle", x0, Wd, 0.0, 0.2, 1.8, iters, post=("shift", -20.0))
    finally:
        ctx.host_unregister(x0); ctx.host_unregister(Wd)
    ctx.host_unregister(x0)                               # idempotent
    assert np.array_equal(a["X_end"], b["X_end"]) and np.array_equal(a["U"], b["U"]) and np.array_equal(a["naccept"], b["naccept"])
    idx = np.arange(0, N, 997)
    ref = _oracle().rollout(_oracle().Net(net.as_tuples()), "Unicycle", 4, 1, 2, x0[idx], Wd[:, idx], 0.0, 0.2, 1.8, iters,
                            post=("shift", -20.0))
    np.testing.assert_allclose(a["X_end"][:, idx], ref["X_end"], rtol=0, atol=TRAJ_ATOL)
