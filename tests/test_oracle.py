"""CPU tests of the oracle itself: golden pins, the VerticalCAS KAT, structural properties."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_net
from oracle import cbind, deepz_np as dz
import vcas_kat

TORA_LO, TORA_HI = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]


def _tables(lo, hi, part):
    cent, rad = dz.split_tables(lo, hi, part)
    return cent, [np.full(len(c), r) for c, r in zip(cent, rad)]


def test_vcas_verdicts_numpy_oracle():
    nets = [load_net(f"VerticalCAS_{i}").as_tuples() for i in range(1, 10)]

    def fh(adv, c, G):
        c2, G2 = dz.deepz_forward(nets[adv], c, G)
        return dz.low_high(c2, G2)[1]

    for hdot0, want in vcas_kat.EXPECTED.items():
        verdict, _ = vcas_kat.run_instance(hdot0, fh)
        assert verdict == want, (hdot0, verdict)


def test_vcas_verdicts_c_oracle():
    nets = [cbind.Net(load_net(f"VerticalCAS_{i}").as_tuples()) for i in range(1, 10)]

    def fh(adv, c, G):
        r = cbind.deepz_zonotopes(nets[adv], c[None, :], [0, G.shape[1]], G.T)
        return r["hi"][0]

    for hdot0, want in vcas_kat.EXPECTED.items():
        verdict, _ = vcas_kat.run_instance(hdot0, fh)
        assert verdict == want, (hdot0, verdict)


def test_tora_golden_numpy_and_c():
    z = np.load(os.path.join(GOLDEN, "tora_oracle.npz"))
    net = load_net("TORA_ReLU").as_tuples()
    cn = cbind.Net(net)
    for tag, part in (("p4435", [4, 4, 3, 5]), ("p1010101", [10, 10, 10, 1])):
        cent, radii = _tables(TORA_LO, TORA_HI, part)
        r = cbind.deepz_boxgrid(cn, part, cent, radii, post=("shift", -10.0))
        np.testing.assert_allclose(r["lo"], z[f"{tag}_lo"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["hi"], z[f"{tag}_hi"], rtol=0, atol=1e-12)
        assert r["stats"]["sum_p_in"] == list(z[f"{tag}_p_in"].sum(0))
    l, h = dz.deepz_boxgrid(net, TORA_LO, TORA_HI, [4, 4, 3, 5], post=("shift", -10.0))
    np.testing.assert_allclose(l, z["p4435_lo"], rtol=0, atol=1e-14)
    # whole-X0 box gives u in [-0.2087, 0.2657]
    assert abs(z["whole_lo"][0] + 0.2087) < 1e-4 and abs(z["whole_hi"][0] - 0.2657) < 1e-4


def test_quadrotor_sigmoid_golden():
    z = np.load(os.path.join(GOLDEN, "quadrotor_oracle.npz"))
    net = load_net("Quadrotor").as_tuples()
    cn = cbind.Net(net)
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    part = [2] * 6 + [1] * 6
    cent, radii = _tables(-r, r, part)
    out = cbind.deepz_boxgrid(cn, part, cent, radii, keep_cap=256)
    np.testing.assert_allclose(out["lo"], z["lo"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["hi"], z["hi"], rtol=0, atol=1e-12)
    p = out["ncols"][0]
    assert p == z["cell0_G"].shape[1] == 6 + 64 * 3
    # 1e-12 relative to the zonotope's scale: entries that are sums with cancellation cannot
    # agree element-wise-relative across different (equally valid) summation orders
    scale = np.abs(z["cell0_G"]).max()
    np.testing.assert_allclose(out["G"][0, :p, :].T, z["cell0_G"], rtol=0, atol=1e-12 * scale)


def test_degenerate_cell_equals_pointwise_forward():
    """A zero-radius cell has no generators: DeepZ must equal pointwise forward (SURVEY 8c KAT 2)."""
    rng = np.random.default_rng(1)
    for name in ("TORA_ReLU", "ACC_tanh", "Quadrotor", "Unicycle"):
        net = load_net(name).as_tuples()
        n = net[0][0].shape[1]
        x = rng.uniform(-1, 1, n)
        c, G = dz.deepz_forward(net, x, np.zeros((n, 0)))
        assert G.shape[1] == 0 or np.all(G == 0)
        np.testing.assert_allclose(c, dz.forward_point(net, x), rtol=1e-13, atol=1e-15)
        cn = cbind.Net(net)
        r = cbind.deepz_zonotopes(cn, x[None, :], [0, 0], np.zeros((0, n)))
        u = cbind.forward_points(cn, x[None, :])
        np.testing.assert_allclose(r["lo"], u, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(r["hi"], u, rtol=1e-13, atol=1e-15)


def test_soundness_by_sampling():
    rng = np.random.default_rng(2)
    for name, post in (("TORA_ReLU", ("shift", -10.0)), ("TORA_sigmoid", ("scale", 11.0)),
                       ("TORA_ReLUtanh", ("scale", 11.0)), ("Unicycle", ("shift", -20.0))):
        net = load_net(name).as_tuples()
        n = net[0][0].shape[1]
        c0 = rng.uniform(-0.5, 0.5, n)
        r0 = rng.uniform(0.01, 0.05, n)
        c, G = dz.hyperrectangle_to_zonotope(c0, r0)
        _, _, lo, hi = dz.controller_forward(net, c, G, None, post)
        xs = c0 + r0 * rng.uniform(-1, 1, (500, n))
        us = np.array([dz.apply_post_point(post, dz.forward_point(net, x)) for x in xs])
        assert np.all(us >= lo - 1e-12) and np.all(us <= hi + 1e-12)


def test_hand_example_relu():
    """2 -> 2 ReLU layer worked by hand: x in [-1,1]^2, y1 = x1 + x2 (unstable), y2 = x1 - x2 + 3 (active)."""
    net = [(np.array([[1.0, 1.0], [1.0, -1.0]]), np.array([0.0, 3.0]), "ReLU")]
    c, G = dz.hyperrectangle_to_zonotope([0.0, 0.0], [1.0, 1.0])
    c2, G2 = dz.deepz_forward(net, c, G)
    # neuron 1: l=-2,u=2 -> lambda=1/2, mu=1/2: c=0.5, row=[.5,.5], new generator 0.5
    np.testing.assert_allclose(c2, [0.5, 3.0])
    np.testing.assert_allclose(G2, [[0.5, 0.5, 0.5], [1.0, -1.0, 0.0]])


def test_id_network_is_exact_affine_map():
    """random_control_problems recipe (src/utils.jl:45-62): Id-only nets are exact affine maps."""
    rng = np.random.default_rng(3)
    W = 1.1 * (rng.random((3, 5)) - 0.5)
    b = 0.1 * (rng.random(3) - 0.5)
    c0, r0 = rng.random(5), rng.random(5)
    c, G = dz.hyperrectangle_to_zonotope(c0, r0)
    c2, G2 = dz.deepz_forward([(W, b, "Id")], c, G)
    np.testing.assert_allclose(c2, W @ c0 + b, rtol=1e-15)
    np.testing.assert_allclose(G2, W @ np.diag(r0), rtol=1e-15)


def test_one_row_layer_collapses_to_single_generator():
    net = load_net("TORA_ReLU").as_tuples()
    c, G = dz.hyperrectangle_to_zonotope([0.65, -0.65, -0.35, 0.55], [0.05] * 4)
    layers = []
    dz.deepz_forward(net, c, G, layers_out=layers)
    assert layers[-1][1].shape[1] <= 2  # one generator (+ one if the output ReLU is unstable)


def test_split_order_dimension_one_fastest():
    cells, rad = dz.split_box([0, 10, 100], [3, 12, 101], [3, 2, 1])
    np.testing.assert_allclose(rad, [0.5, 0.5, 0.5])
    np.testing.assert_allclose(cells[:, 0], [0.5, 1.5, 2.5, 0.5, 1.5, 2.5])
    np.testing.assert_allclose(cells[:, 1], [10.5, 10.5, 10.5, 11.5, 11.5, 11.5])
    for i in range(6):
        idx = dz.cell_multi_index(i, [3, 2, 1])
        assert idx == [i % 3, i // 3, 0]
    with pytest.raises(ValueError):
        dz.split_tables([0], [1], [0])


def test_julia_range_rationalisation():
    # 0.605:0.01 -> exact decimals k/200, e.g. element 2 is the double nearest 0.625
    r = dz.julia_range(0.605, 0.01, 10)
    assert r[2] == 0.625 and r[0] == 0.605 and r[9] == 0.695


def test_c_oracle_threads_and_flops():
    assert cbind.num_threads() >= 1
    net = load_net("TORA_ReLU").as_tuples()
    cn = cbind.Net(net)
    cent, radii = _tables(TORA_LO, TORA_HI, [10, 10, 10, 1])
    r = cbind.deepz_boxgrid(cn, [10, 10, 10, 1], cent, radii, post=("shift", -10.0))
    # mean 0.327 MFLOP per cell on this workload
    assert abs(r["stats"]["flops"] / 1000 / 1e6 - 0.327) < 0.002


# ---- order reduction of the output zonotope (TaylorModelReconstructor(box=false), src/setops.jl:84-101) ----
def test_reduce_order_hand_example():
    # 2-dim zonotope with 5 generators; weights |g|_1 - |g|_inf = [0, 1, 0.5, 0, 0.25]
    c = np.array([1.0, -1.0])
    G = np.array([[2.0, 1.0, 0.5, 0.0, -0.25],
                  [0.0, 3.0, -4.0, 7.0, 1.0]])
    c2, R = dz.reduce_order_gir05(c, G, 2)
    assert R.shape == (2, 4) and np.array_equal(c2, c)
    # kept: generators 2 and 3 (weights 1 and 0.5); boxed in sorted order: 5 (0.25), then 1 and 4 (ties keep their order)
    assert np.array_equal(R[:, :2], G[:, [1, 2]])
    assert np.array_equal(R[:, 2:], np.diag([0.25 + 2.0 + 0.0, 1.0 + 0.0 + 7.0]))
    # the layout src/setops.jl:96-101 asserts: (m, 2m) with a diagonal tail
    D = R[:, 2:]
    assert np.count_nonzero(D - np.diag(np.diag(D))) == 0
    # order 1 = the interval hull, no sorting
    _, B = dz.reduce_order_gir05(c, G, 1)
    assert np.array_equal(B, np.diag(np.abs(G).sum(axis=1)))
    with pytest.raises(IndexError):
        dz.reduce_order_gir05(c, G[:, :1], 2)


@pytest.mark.parametrize("m,order", [(2, 2), (3, 2), (3, 3), (6, 2), (3, 1)])
def test_reduce_order_numpy_vs_c_and_box_invariance(m, order):
    rng = np.random.default_rng(100 * m + order)
    N, cap = 40, 64
    ncols = rng.integers(m * (order - 1), cap + 1, size=N).astype(np.int32)
    c = rng.normal(size=(N, m))
    G = np.zeros((N, cap, m))
    for i in range(N):
        G[i, :ncols[i]] = rng.normal(size=(ncols[i], m)) * rng.choice([1.0, 1e-3, 0.0], size=(ncols[i], 1), p=[.7, .2, .1])
        if ncols[i] > 3:
            G[i, 2] = G[i, 0]                      # equal weights: the stable order decides
    oc, oG = cbind.reduce_order(ncols, c, G, order)
    for i in range(N):
        p = ncols[i]
        rc, rG = dz.reduce_order_gir05(c[i], G[i, :p].T, order)
        assert np.array_equal(oc[i], rc)
        assert np.array_equal(oG[i].T, rG)          # same sums in the same order: bit-identical
        # the reduction keeps the box of the zonotope (up to the order of the sums)
        lo0, hi0 = dz.low_high(c[i], G[i, :p].T)
        lo1, hi1 = dz.low_high(rc, rG)
        np.testing.assert_allclose(lo1, lo0, rtol=0, atol=1e-12)
        np.testing.assert_allclose(hi1, hi0, rtol=0, atol=1e-12)
    with pytest.raises(IndexError):
        if order > 1:
            cbind.reduce_order(np.zeros(1, dtype=np.int32), c[:1], G[:1], order)
        else:
            raise IndexError


# ---- _project_oa of Taylor-model reach-sets (src/setops.jl:9-12) ----
def _random_taylor_models(rng, N, n, order_t, order_q, scale=0.1):
    exps = dz.monomial_table(n, order_q)
    M = exps.shape[0]
    deg = exps.sum(axis=1)
    coeffs = rng.normal(size=(N, n, order_t + 1, M)) * (scale ** deg)[None, None, None, :] / (1.0 + np.arange(order_t + 1))[None, None, :, None]
    for v in range(n):                                     # time-0 term: identity map of the initial box plus an offset
        coeffs[:, v, 0, :] = 0.0
        coeffs[:, v, 0, 0] = rng.normal(size=N)
        coeffs[:, v, 0, 1 + v] = scale
    rem = np.sort(rng.normal(size=(N, n, 2)) * 1e-6, axis=2)
    dt = rng.uniform(0.05, 0.2, size=N)
    return exps, coeffs, rem, dt


def test_project_oa_hand_example():
    # x(t) = (1 + t) + (0.5 + t) xi1 + 0.25 xi1^2 - 0.125 xi1 xi2 + [-0.01, 0.02];  y(t) = 2 + xi2  (no remainder)
    exps = dz.monomial_table(2, 2)                         # [0,0],[1,0],[0,1],[2,0],[1,1],[0,2]
    coeffs = np.zeros((2, 2, 6))
    coeffs[0, 0] = [1.0, 0.5, 0.0, 0.25, -0.125, 0.0]; coeffs[0, 1] = [1.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    coeffs[1, 0] = [2.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    rem = np.array([[-0.01, 0.02], [0.0, 0.0]])
    c, G = dz.project_oa_taylor_model(exps, coeffs, rem, 0.5, [0, 1])
    # at t = 0.5: constant 1.5, linear 1.0 xi1; nonlinear range 0.25 [0,1] + 0.125 [-1,1] = [-0.125, 0.375]; with the
    # remainder [-0.135, 0.395]: centre 1.5 + 0.13, radius 0.265
    np.testing.assert_allclose(c, [1.63, 2.0], rtol=0, atol=1e-14)
    assert G.shape == (2, 3)                               # columns: xi1, xi2, remainder of x (y has none: zero column removed)
    np.testing.assert_allclose(G, [[1.0, 0.0, 0.265], [0.0, 1.0, 0.0]], rtol=0, atol=1e-14)
    c1, G1 = dz.project_oa_taylor_model(exps, coeffs, rem, 0.5, [1], remove_zero_generators=False)
    assert G1.shape == (1, 4) and np.array_equal(G1[0], [0.0, 1.0, 0.0, 0.0]) and c1[0] == 2.0


def test_remove_redundant_generators_hand_examples():
    """LazySets.remove_redundant_generators (the default of overapproximate(::Vector{TaylorModel1}, Zonotope), which
    _project_oa runs, src/setops.jl:10): zero columns dropped, multiples merged into the first of them."""
    G = np.array([[1.0, 2.0, 0.0, 0.0], [0.0, 0.0, 0.0, 3.0]])
    np.testing.assert_array_equal(dz.remove_redundant_generators(G), [[3.0, 0.0], [0.0, 3.0]])
    G = np.array([[1.0, -2.0, 0.5], [1.0, -2.0, 0.25]])                   # negative factor: subtracted
    np.testing.assert_array_equal(dz.remove_redundant_generators(G), [[3.0, 0.5], [3.0, 0.25]])
    G = np.array([[1.0, 1.0, 1.0], [0.0, 1e-20, 0.0]])                    # 1e-20 is approximately zero: all three are multiples
    np.testing.assert_array_equal(dz.remove_redundant_generators(G), [[3.0], [1e-20]])
    assert dz.ismultiple([1.0, 2.0], [2.0, 4.0]) == (True, 0.5)
    assert dz.ismultiple([0.0, 0.0], [0.0, 0.0]) == (True, 0.0)
    assert dz.ismultiple([1.0, 0.0], [1.0, 1.0])[0] is False
    # the box of the zonotope does not change, the set neither (generators parallel to each other add up)
    rng = np.random.default_rng(0)
    G = rng.normal(size=(4, 9)); G[:, 5] = -0.3 * G[:, 1]; G[:, 7] = 2.0 * G[:, 1]; G[:, 3] = 0.0
    R = dz.remove_redundant_generators(G)
    assert R.shape == (4, 6)
    np.testing.assert_allclose(np.abs(R).sum(axis=1), np.abs(G).sum(axis=1), rtol=1e-14)


def test_project_oa_with_redundant_generators():
    # x(t) = 1 + xi1 + [-0.1, 0.1];  u(t) = 0.5 xi2 + [-0.2, 0.2]: each linear column is parallel to the remainder generator of
    # its own component, so overapproximate's default merges them: G = diag(1.1, 0.7) instead of four columns
    exps = dz.monomial_table(2, 1)
    coeffs = np.zeros((2, 1, 3)); coeffs[0, 0] = [1.0, 1.0, 0.0]; coeffs[1, 0] = [0.0, 0.0, 0.5]
    rem = np.array([[-0.1, 0.1], [-0.2, 0.2]])
    c, G = dz.project_oa_taylor_model(exps, coeffs, rem, 0.0, [0, 1], remove_redundant=True)
    np.testing.assert_allclose(G, [[1.1, 0.0], [0.0, 0.7]], rtol=0, atol=1e-15)
    c2, G2 = dz.project_oa_taylor_model(exps, coeffs, rem, 0.0, [0, 1], remove_redundant=False)
    assert G2.shape == (2, 4) and np.array_equal(c, c2)
    # projection AFTER the merge: keeping only x leaves its merged generator
    c3, G3 = dz.project_oa_taylor_model(exps, coeffs, rem, 0.0, [0], remove_redundant=True)
    np.testing.assert_allclose(G3, [[1.1]], rtol=0, atol=1e-15)


def test_project_oa_encloses_the_taylor_model():
    rng = np.random.default_rng(11)
    exps, coeffs, rem, dt = _random_taylor_models(rng, 5, 3, 4, 3)
    for i in range(5):
        c, G = dz.project_oa_taylor_model(exps, coeffs[i], rem[i], dt[i], [0, 2])
        lo, hi = dz.low_high(c, G)
        xi = rng.uniform(-1, 1, size=(200, 3)); xi[:8] = np.array(list(__import__("itertools").product([-1, 1], repeat=3)))
        for x in xi:
            val = dz.eval_taylor_model_point(exps, coeffs[i], dt[i], x)[[0, 2]]
            assert np.all(val + rem[i][[0, 2], 0] >= lo - 1e-12) and np.all(val + rem[i][[0, 2], 1] <= hi + 1e-12)


def test_setops_golden_fixture():
    """Regression pins of the 8f restatements (tests/golden/setops_oracle.npz, written by make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "setops_oracle.npz"))
    q = np.load(os.path.join(GOLDEN, "quadrotor_oracle.npz"))
    ncols = np.array([q["cell0_G"].shape[1]], dtype=np.int32)
    for order in (1, 2, 3):
        rc, rG = dz.reduce_order_gir05(q["cell0_c"], q["cell0_G"], order)
        assert np.array_equal(rG, z[f"red{order}_G"]) and np.array_equal(rc, z[f"red{order}_c"])
        oc, oG = cbind.reduce_order(ncols, q["cell0_c"][None, :], q["cell0_G"].T[None, :, :], order)
        assert np.array_equal(oG[0].T, z[f"red{order}_G"])
    for i in range(z["tm_dt"].shape[0]):
        pc, pG = dz.project_oa_taylor_model(z["tm_exps"], z["tm_coeffs"][i], z["tm_rem"][i], z["tm_dt"][i], [0, 2])
        assert np.array_equal(pc, z[f"proj{i}_c"]) and np.array_equal(pG, z[f"proj{i}_G"])
