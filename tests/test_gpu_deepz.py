"""GPU parity tests of the DeepZ path: libclr_b200 (through the C ABI) against the CPU oracle and
the committed golden fixtures.  Tolerances (BASELINE.json north_star): output boxes 1e-10 absolute,
centres / generators 1e-12 relative to the zonotope's scale, cell ordering bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_net
import vcas_kat

pytestmark = pytest.mark.gpu

BOX_ATOL = 1e-10
TORA_LO, TORA_HI = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]


@pytest.fixture(scope="module")
def ctx():
    from clr_b200 import runtime
    c = runtime.Context(0)
    yield c
    c.close()


def _oracle():
    from oracle import cbind
    return cbind


def _cmp_boxes(got, ref, atol=BOX_ATOL):
    np.testing.assert_allclose(got["lo"], ref["lo"], rtol=0, atol=atol)
    np.testing.assert_allclose(got["hi"], ref["hi"], rtol=0, atol=atol)


def test_tora_boxgrid_golden_and_oracle(ctx):
    import clr_b200
    z = np.load(os.path.join(GOLDEN, "tora_oracle.npz"))
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    on = _oracle().Net(net.as_tuples())
    for tag, part in (("p4435", [4, 4, 3, 5]), ("p1010101", [10, 10, 10, 1])):
        g = clr_b200.split_box(TORA_LO, TORA_HI, part)
        got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, post=("shift", -10.0), hull=True)
        np.testing.assert_allclose(got["lo"], z[f"{tag}_lo"], rtol=0, atol=BOX_ATOL)
        np.testing.assert_allclose(got["hi"], z[f"{tag}_hi"], rtol=0, atol=BOX_ATOL)
        ref = _oracle().deepz_boxgrid(on, g.partition, g.centers, g.radii, post=("shift", -10.0))
        _cmp_boxes(got, ref)
        assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]
        assert got["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"]
        assert got["stats"]["flops"] == ref["stats"]["flops"]
        assert got["hull_lo"][0] == got["lo"].min() and got["hull_hi"][0] == got["hi"].max()


def test_quadrotor_sigmoid_zonotope_fetch(ctx):
    import clr_b200
    z = np.load(os.path.join(GOLDEN, "quadrotor_oracle.npz"))
    net = load_net("Quadrotor")
    dn = ctx.upload(net)
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    g = clr_b200.split_box(-r, r, [2] * 6 + [1] * 6)
    got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, keep_cap=256)
    np.testing.assert_allclose(got["lo"], z["lo"], rtol=0, atol=BOX_ATOL)
    np.testing.assert_allclose(got["hi"], z["hi"], rtol=0, atol=BOX_ATOL)
    p = got["ncols"][0]
    assert p == z["cell0_G"].shape[1] == 6 + 64 * 3
    scale = np.abs(z["cell0_G"]).max()
    np.testing.assert_allclose(got["G"][0, :p, :].T, z["cell0_G"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(got["c"][0], z["cell0_c"], rtol=1e-12, atol=1e-14)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii, keep_cap=256)
    assert np.array_equal(got["ncols"], ref["ncols"])
    np.testing.assert_allclose(got["G"], ref["G"], rtol=0, atol=1e-12 * scale)


def test_vcas_verdicts(ctx):
    nets = [ctx.upload(load_net(f"VerticalCAS_{i}")) for i in range(1, 10)]

    def fh(adv, c, G):
        r = ctx.deepz_zonotopes(nets[adv], c[None, :], [0, G.shape[1]], G.T)
        return r["hi"][0]

    for hdot0, want in vcas_kat.EXPECTED.items():
        verdict, _ = vcas_kat.run_instance(hdot0, fh)
        assert verdict == want, (hdot0, verdict)


@pytest.mark.parametrize("name,post", [("TORA_ReLU", ("shift", -10.0)), ("TORA_sigmoid", ("scale", 11.0)),
                                       ("TORA_ReLUtanh", ("scale", 11.0)), ("Unicycle", ("shift", -20.0)),
                                       ("AttitudeControl", None), ("InvertedPendulum", None)])
def test_random_zonotopes_vs_oracle(ctx, name, post):
    from clr_b200 import workloads
    net = load_net(name)
    n = net.dim_in
    rng = np.random.default_rng(11)
    C, off, G = workloads.random_zonotopes(rng, 300, n, 0, 14, -0.6, 0.6, 0.02)
    got = ctx.deepz_zonotopes(ctx.upload(net), C, off, G, post=post, keep_cap=600)
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), C, off, G, post=post, keep_cap=600)
    _cmp_boxes(got, ref)
    assert np.array_equal(got["ncols"], ref["ncols"])
    scale = max(np.abs(ref["G"]).max(), 1e-300)
    np.testing.assert_allclose(got["G"], ref["G"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(got["c"], ref["c"], rtol=0, atol=1e-12 * max(np.abs(ref["c"]).max(), scale))
    assert got["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"]
    assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]


def test_acc_affine_preprocessing(ctx):
    """ACC: x -> [30, 1.4, M x] (models/ACC/ACC.jl:81-99) folded in as an affine pre-layer."""
    net = load_net("ACC_relu")
    A = np.zeros((5, 6))
    A[2, 4] = 1.0                      # v_ego
    A[3, 0], A[3, 3] = 1.0, -1.0       # D_rel = x_lead - x_ego
    A[4, 1], A[4, 4] = 1.0, -1.0       # v_rel
    d = np.array([30.0, 1.4, 0.0, 0.0, 0.0])
    rng = np.random.default_rng(5)
    from clr_b200 import workloads
    C, off, G = workloads.random_zonotopes(rng, 200, 6, 2, 14, 0.0, 1.0, 0.05)
    C += np.array([90.0, 32.0, 0.0, 10.0, 30.0, 0.0])
    got = ctx.deepz_zonotopes(ctx.upload(net), C, off, G, pre=(A, d))
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), C, off, G, pre=(A, d))
    _cmp_boxes(got, ref)
    assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]


@pytest.mark.parametrize("post", [("matrix", np.array([[1.0, -2.0, 0.5], [0.0, 1.0, 1.0]])),
                                  ("matrix", np.array([[0.3, -0.2, 1.5]])),
                                  ("project", [2, 0]), ("project", [1])])
def test_matrix_and_projection_postprocessing(ctx, post):
    import clr_b200
    net = load_net("Quadrotor")
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    g = clr_b200.split_box(-r, r, [2, 2, 2] + [1] * 9)
    got = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii, post=post, keep_cap=256)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii, post=post,
                                  keep_cap=256)
    _cmp_boxes(got, ref)
    assert np.array_equal(got["ncols"], ref["ncols"])
    # 1e-12 relative to the zonotope's scale (largest |entry| of centre and generators): entries that are
    # sums with cancellation cannot agree element-wise-relative across equally valid summation orders
    scale = max(np.abs(ref["c"]).max(), np.abs(ref["G"]).max())
    np.testing.assert_allclose(got["G"], ref["G"], rtol=0, atol=1e-12 * scale)          # signed


@pytest.mark.parametrize("regime,count", [("stable", 3000), ("mixed", 600), ("dense", 160)])
def test_c4_synthetic_regimes(ctx, regime, count):
    """6 -> 100^4 -> 1 ReLU net of BASELINE's scaling config: small tiles, overflow retries and
    (dense regime) the global-store pass for cells that do not fit in shared memory."""
    from clr_b200 import workloads
    net = workloads.synthetic_relu_net()
    g = workloads.c4_grid(regime, last_dim_blocks=1)
    begin = 12345
    got = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii, cell_begin=begin, cell_count=count,
                            hull=True)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii,
                                  cell_begin=begin, cell_count=count)
    scale = max(1.0, np.abs(ref["hi"]).max())
    _cmp_boxes(got, ref, atol=BOX_ATOL * scale)
    assert got["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"]
    assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]
    assert got["hull_lo"][0] == got["lo"].min() and got["hull_hi"][0] == got["hi"].max()


def test_cut_point_partition_and_flat_dims(ctx):
    """Cut-point BoxSplitter (models/InvertedPendulum/InvertedPendulum.jl:210) and zero-radius blocks."""
    import clr_b200
    net = load_net("InvertedPendulum")
    g = clr_b200.split_box([1.0, 0.0], [1.2, 0.2], [[1.1, 1.16], [0.09, 0.145, 0.18]])
    assert len(g) == 12
    got = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii)
    _cmp_boxes(got, ref)
    g2 = clr_b200.split_box([1.0, 0.1], [1.2, 0.1], [4, 1])          # second dimension is flat
    got = ctx.deepz_boxgrid(ctx.upload(net), g2.partition, g2.centers, g2.radii)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g2.partition, g2.centers, g2.radii)
    _cmp_boxes(got, ref)
    assert got["stats"]["sum_p_in"][0] == 4


def test_degenerate_cells_equal_pointwise_forward(ctx):
    rng = np.random.default_rng(1)
    for name in ("TORA_ReLU", "ACC_tanh", "Quadrotor", "Unicycle"):
        net = load_net(name)
        dn = ctx.upload(net)
        X = rng.uniform(-1, 1, (50, net.dim_in))
        r = ctx.deepz_zonotopes(dn, X, np.zeros(51, dtype=np.int64), np.zeros((0, net.dim_in)))
        u = ctx.forward_points(dn, X)
        np.testing.assert_allclose(r["lo"], u, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(r["hi"], u, rtol=1e-12, atol=1e-14)
        # the padded layout with no generator capacity at all, and an empty batch
        rp = ctx.deepz_zonotopes_padded(dn, X, np.zeros((50, 0, net.dim_in)), np.zeros(50, dtype=np.int32))
        assert np.array_equal(rp["lo"], r["lo"]) and np.array_equal(rp["hi"], r["hi"])
        re = ctx.deepz_zonotopes_padded(dn, X[:0], np.zeros((0, 3, net.dim_in)), np.zeros(0, dtype=np.int32))
        assert re["lo"].shape == (0, r["lo"].shape[1])


def test_shards_concatenate_to_the_full_result(ctx):
    import clr_b200
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    g = clr_b200.split_box(TORA_LO, TORA_HI, [10, 10, 10, 1])
    full = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, post=("shift", -10.0))
    parts = []
    for rank in range(3):
        b, cnt = g.shard(rank, 3)
        parts.append(ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=b, cell_count=cnt,
                                       post=("shift", -10.0))["lo"])
    assert np.array_equal(np.concatenate(parts), full["lo"])      # bit-exact, ordering included
    ctx.set_tile_cells(1)
    one = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, post=("shift", -10.0))
    ctx.set_tile_cells(0)
    assert np.array_equal(one["lo"], full["lo"]) and np.array_equal(one["hi"], full["hi"])


def test_empty_and_error_paths(ctx):
    import clr_b200
    from clr_b200 import runtime
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    g = clr_b200.split_box(TORA_LO, TORA_HI, [2, 2, 2, 2])
    out = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=3, cell_count=0)
    assert out["lo"].shape == (0, 1)
    with pytest.raises(runtime.ClrArgumentError):
        ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=10, cell_count=10)
    g3 = clr_b200.split_box([0, 0, 0], [1, 1, 1], [2, 2, 2])
    with pytest.raises(runtime.ClrArgumentError):
        ctx.deepz_boxgrid(dn, g3.partition, g3.centers, g3.radii)       # 3-dim cells into a 4-input net
    ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii)
    with pytest.raises(runtime.ClrError):
        ctx.fetch_zonotopes(16, 1, 4)                                    # nothing was kept by the last run
    quad = load_net("Quadrotor")
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    gq = clr_b200.split_box(-r, r, [2] + [1] * 11)
    with pytest.raises(runtime.ClrError):
        ctx.deepz_boxgrid(ctx.upload(quad), gq.partition, gq.centers, gq.radii, keep_cap=4)   # 198 generators


def test_full_size_properties(ctx):
    """BASELINE size (10^6 cells, 6 -> 100^4 -> 1): soundness by sampling, shard consistency and
    monotonicity of the hull, none of which needs the oracle at this size."""
    from clr_b200 import workloads
    net = workloads.synthetic_relu_net()
    dn = ctx.upload(net)
    g = workloads.c4_grid("stable")
    assert len(g) == 10 ** 6
    out = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, hull=True)
    assert np.all(out["lo"] <= out["hi"])
    assert out["hull_lo"][0] == out["lo"].min() and out["hull_hi"][0] == out["hi"].max()
    rng = np.random.default_rng(7)
    ids = rng.integers(0, len(g), 2000)
    X = np.empty((2000, 6))
    for k, i in enumerate(ids):
        c, r = g.cell(int(i))
        X[k] = c + r * rng.uniform(-1, 1, 6)
    U = ctx.forward_points(dn, X)
    assert np.all(U[:, 0] >= out["lo"][ids, 0] - 1e-9) and np.all(U[:, 0] <= out["hi"][ids, 0] + 1e-9)
    b, cnt = g.shard(5, 8)
    part = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=b, cell_count=cnt)
    assert np.array_equal(part["lo"], out["lo"][b:b + cnt]) and np.array_equal(part["hi"], out["hi"][b:b + cnt])
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii,
                                  cell_begin=b, cell_count=2000)
    np.testing.assert_allclose(part["lo"][:2000], ref["lo"], rtol=0, atol=BOX_ATOL)


def test_quadrotor_c3_4096_cells(ctx):
    """BASELINE config C3: Quadrotor controller (12 -> 64^3 -> 3, sigmoid), X0 of models/Quadrotor/Quadrotor.jl:98-99
    split 4^6 (4096 cells, 198 generators per output zonotope).  Blocks of cells are checked against the oracle,
    all cells through soundness and the hull."""
    import clr_b200
    net = load_net("Quadrotor")
    dn = ctx.upload(net)
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    g = clr_b200.split_box(-r, r, [4] * 6 + [1] * 6)
    assert len(g) == 4096
    out = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, hull=True)
    assert out["stats"]["max_p_out"] == 6 + 3 * 64 and out["stats"]["sum_unstable"] == [4096 * 64] * 3 + [0]
    on = _oracle().Net(net.as_tuples())
    for b in range(0, 4096, 512):
        ref = _oracle().deepz_boxgrid(on, g.partition, g.centers, g.radii, cell_begin=b, cell_count=64)
        np.testing.assert_allclose(out["lo"][b:b + 64], ref["lo"], rtol=0, atol=BOX_ATOL)
        np.testing.assert_allclose(out["hi"][b:b + 64], ref["hi"], rtol=0, atol=BOX_ATOL)
    rng = np.random.default_rng(3)
    ids = rng.integers(0, 4096, 500)
    X = np.array([g.cell(int(i))[0] + g.cell(int(i))[1] * rng.uniform(-1, 1, 12) for i in ids])
    U = ctx.forward_points(dn, X)
    assert np.all(U >= out["lo"][ids] - 1e-12) and np.all(U <= out["hi"][ids] + 1e-12)
    np.testing.assert_array_equal(out["hull_lo"], out["lo"].min(0))


def test_tora_c2_ten_control_periods(ctx):
    """BASELINE config C2: TORA, 10^3 cells per step, 10 control periods: k = 1 is the box split of X0, k > 1 are
    10^3 synthetic zonotopes per step (n = 4, p0 = 10)."""
    import clr_b200
    from clr_b200 import workloads
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    on = _oracle().Net(net.as_tuples())
    g = clr_b200.split_box(TORA_LO, TORA_HI, [10, 10, 10, 1])
    got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, post=("shift", -10.0))
    ref = _oracle().deepz_boxgrid(on, g.partition, g.centers, g.radii, post=("shift", -10.0))
    _cmp_boxes(got, ref)
    rng = np.random.default_rng(12)
    for k in range(2, 11):
        C, off, G = workloads.random_zonotopes(rng, 1000, 4, 10, 10, -0.7, 0.7, 0.9 ** k * 0.02)
        got = ctx.deepz_zonotopes(dn, C, off, G, post=("shift", -10.0))
        ref = _oracle().deepz_zonotopes(on, C, off, G, post=("shift", -10.0))
        _cmp_boxes(got, ref)
        assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]


def test_zonotope_splitter_cells(ctx):
    """ZonotopeSplitter cells (src/splitters.jl:30-43) go through clr_deepz_zonotopes like any other cell batch."""
    import clr_b200
    net = load_net("TORA_ReLU")
    rng = np.random.default_rng(9)
    c = np.array([0.65, -0.65, -0.35, 0.55])
    G = np.hstack([np.diag([0.05] * 4), 1e-3 * rng.uniform(-1, 1, (4, 3))])
    cells = clr_b200.ZonotopeSplitter([0, 1, 2, 4], [2, 2, 1, 1]).apply(c, G)
    assert len(cells) == 64
    got = ctx.deepz_zonotopes(ctx.upload(net), cells.C, cells.col_off, cells.G, post=("shift", -10.0), hull=True)
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), cells.C, cells.col_off, cells.G, post=("shift", -10.0))
    _cmp_boxes(got, ref)
    assert got["hull_lo"][0] == got["lo"].min() and got["hull_hi"][0] == got["hi"].max()


@pytest.mark.parametrize("regime,count", [("mixed", 400), ("stable", 1500)])
def test_results_do_not_depend_on_the_tiling(ctx, regime, count):
    """The in-place layer products move columns while other warps still read neighbours: every tile size must give
    bit-identical boxes (and the oracle's, within tolerance)."""
    from clr_b200 import workloads
    net = workloads.synthetic_relu_net()
    dn = ctx.upload(net)
    g = workloads.c4_grid(regime, last_dim_blocks=1)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii, cell_begin=777,
                                  cell_count=count)
    base = None
    for T in (0, 1, 2, 3, 5, 8, 13, 21, 34):
        ctx.set_tile_cells(T)
        got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=777, cell_count=count, stats=False)
        if base is None:
            base = got
            _cmp_boxes(got, ref, atol=BOX_ATOL * max(1.0, np.abs(ref["hi"]).max()))
        else:
            assert np.array_equal(got["lo"], base["lo"]) and np.array_equal(got["hi"], base["hi"]), T
    ctx.set_tile_cells(0)


def test_device_resident_io_and_two_contexts(ctx):
    """CLR_FLAG_DEVICE_IO: inputs and outputs as device pointers (torch CUDA tensors), for zonotope batches and box
    grids, on a second context of the same device; results equal the host-buffer path bit for bit."""
    import torch
    import clr_b200
    from clr_b200 import _lib, runtime, workloads
    net = load_net("TORA_ReLU")
    rng = np.random.default_rng(31)
    C, off, G = workloads.random_zonotopes(rng, 500, 4, 0, 12, -0.6, 0.6, 0.02)
    host = ctx.deepz_zonotopes(ctx.upload(net), C, off, G, post=("shift", -10.0), hull=True)
    ctx2 = runtime.Context(0)
    try:
        dn2 = ctx2.upload(net)
        dev = torch.device("cuda", 0)
        dC, dO, dG = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (C, off, G))
        lo = torch.empty((500, 1), dtype=torch.float64, device=dev)
        hi = torch.empty_like(lo)
        hl = torch.empty(1, dtype=torch.float64, device=dev)
        hh = torch.empty_like(hl)
        pp = runtime.PrePostArg(None, ("shift", -10.0))
        st = _lib.Stats()
        ctx2.deepz_zonotopes_dev(dn2, 4, 500, dC, dO, dG, pp, lo, hi, hl, hh, stats=st)
        ctx2.synchronize()
        assert np.array_equal(lo.cpu().numpy(), host["lo"]) and np.array_equal(hi.cpu().numpy(), host["hi"])
        assert hl.item() == host["hull_lo"][0] and hh.item() == host["hull_hi"][0]
        assert list(st.sum_p_in)[:4] == host["stats"]["sum_p_in"]
        g = clr_b200.split_box(TORA_LO, TORA_HI, [4, 4, 3, 5])
        cc = torch.from_numpy(np.concatenate(g.centers)).to(dev)
        rr = torch.from_numpy(np.concatenate(g.radii)).to(dev)
        lo2 = torch.empty((240, 1), dtype=torch.float64, device=dev)
        hi2 = torch.empty_like(lo2)
        ctx2.deepz_boxgrid_dev(dn2, 4, g.partition, cc, rr, 0, 240, pp, lo2, hi2)
        ctx2.synchronize()
        ref = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii, post=("shift", -10.0))
        assert np.array_equal(lo2.cpu().numpy(), ref["lo"]) and np.array_equal(hi2.cpu().numpy(), ref["hi"])
        assert ctx2.launch_count > 0 and ctx2.stream != ctx.stream
    finally:
        ctx2.close()


def test_random_networks_fuzz(ctx):
    """Differential fuzz (random_control_problems recipe, src/utils.jl:34-90, widened): random depths, widths,
    activations, pre- and post-processing, box and zonotope cells -- every combination against the oracle."""
    import clr_b200
    from clr_b200 import workloads
    rng = np.random.default_rng(2024)
    acts = ["Id", "ReLU", "Sigmoid", "Tanh"]
    for trial in range(int(os.environ.get("CLR_FUZZ_TRIALS", "36"))):
        depth = int(rng.integers(1, 6))
        n_in = int(rng.integers(1, 9))
        dims = [n_in] + [int(rng.choice([1, 2, 3, 7, 8, 9, 16, 20, 33, 64, 100, 130])) for _ in range(depth)]
        layers = []
        for l in range(depth):
            W = rng.uniform(-1, 1, (dims[l + 1], dims[l])) * np.sqrt(3.0 / dims[l])
            b = rng.uniform(-0.3, 0.3, dims[l + 1])
            layers.append((W, b, str(rng.choice(acts))))
        net = clr_b200.FeedforwardNetwork.from_tuples(layers)
        pre = None
        n0 = n_in
        if trial % 3 == 1:
            n0 = int(rng.integers(1, 7))
            pre = (rng.uniform(-1, 1, (n_in, n0)), rng.uniform(-1, 1, n_in))
        m = dims[-1]
        post = [None, ("shift", 0.75), ("scale", -1.5), ("matrix", rng.uniform(-1, 1, (int(rng.integers(1, 4)), m))),
                ("project", sorted(rng.choice(m, size=min(m, 2), replace=False).tolist()))][trial % 5]
        dn = ctx.upload(net)
        on = _oracle().Net(net.as_tuples())
        if trial % 2 == 0:
            lo = rng.uniform(-1, 0, n0)
            hi = lo + rng.uniform(0.0, 0.3, n0) * (rng.random(n0) > 0.2)        # some flat dimensions
            part = [int(rng.integers(1, 4)) for _ in range(n0)]
            g = clr_b200.split_box(lo, hi, part)
            got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, pre=pre, post=post, keep_cap=1200, hull=True)
            ref = _oracle().deepz_boxgrid(on, g.partition, g.centers, g.radii, pre=pre, post=post, keep_cap=1200)
        else:
            C, off, G = workloads.random_zonotopes(rng, 40, n0, 0, 9, -0.8, 0.8, 0.05)
            got = ctx.deepz_zonotopes(dn, C, off, G, pre=pre, post=post, keep_cap=1200, hull=True)
            ref = _oracle().deepz_zonotopes(on, C, off, G, pre=pre, post=post, keep_cap=1200)
        scale = max(1.0, np.abs(ref["hi"]).max(), np.abs(ref["lo"]).max())
        info = (trial, dims, [a for (_, _, a) in layers], pre is not None, post and post[0])
        np.testing.assert_allclose(got["lo"], ref["lo"], rtol=0, atol=BOX_ATOL * scale, err_msg=str(info))
        np.testing.assert_allclose(got["hi"], ref["hi"], rtol=0, atol=BOX_ATOL * scale, err_msg=str(info))
        assert np.array_equal(got["ncols"], ref["ncols"]), info
        assert got["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"], info
        assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"], info
        gs = max(np.abs(ref["G"]).max(), np.abs(ref["c"]).max(), 1e-300)
        np.testing.assert_allclose(got["G"], ref["G"], rtol=0, atol=1e-12 * gs, err_msg=str(info))       # signed
        np.testing.assert_allclose(got["c"], ref["c"], rtol=0, atol=1e-12 * gs, err_msg=str(info))
        # the same cells as a box-only call (small-call path, lean / no-statistics instantiations, post-processing applied
        # in the output stage): same boxes within the tolerance, hull = min / max of the cells
        st = trial % 4 < 2
        if trial % 2 == 0:
            box = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, pre=pre, post=post, hull=True, stats=st)
        else:
            box = ctx.deepz_zonotopes(dn, C, off, G, pre=pre, post=post, hull=True, stats=st)
        np.testing.assert_allclose(box["lo"], ref["lo"], rtol=0, atol=BOX_ATOL * scale, err_msg=str(info))
        np.testing.assert_allclose(box["hi"], ref["hi"], rtol=0, atol=BOX_ATOL * scale, err_msg=str(info))
        np.testing.assert_array_equal(box["hull_lo"], box["lo"].min(0))
        np.testing.assert_array_equal(box["hull_hi"], box["hi"].max(0))
        if st:
            assert box["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"], info
            assert box["stats"]["max_p_out"] == ref["stats"]["max_p_out"], info


def test_order2_reduction_of_kept_zonotopes(ctx):
    """TaylorModelReconstructor(box=false): Z0 = _reduce_order(U0, 2; force_reduction=true), src/setops.jl:84-86,
    for every cell of the Quadrotor split (3 outputs, 198 generators each) in one call."""
    import clr_b200
    from oracle import deepz_np as dz
    net = load_net("Quadrotor")
    dn = ctx.upload(net)
    r = 0.01 * np.array([0.4] * 6 + [0.0] * 6)
    g = clr_b200.split_box(-r, r, [2] * 6 + [1] * 6)
    got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, keep_cap=256)
    N, m = got["lo"].shape
    for order in (2, 1, 3):
        red = ctx.fetch_reduced(N, m, order)
        assert red["G"].shape == (N, m * order, m)
        # (a) against the oracle's reduction of the very same zonotopes: the same sums in the same order -> bit-identical
        oc, oG = _oracle().reduce_order(got["ncols"], got["c"], got["G"], order)
        assert np.array_equal(red["c"], oc)
        assert np.array_equal(red["G"], oG)
        # (b) against the NumPy restatement on one cell, and the diagonal tail src/setops.jl:96-101 asserts
        p = got["ncols"][5]
        _, rG = dz.reduce_order_gir05(got["c"][5], got["G"][5, :p].T, order)
        assert np.array_equal(red["G"][5].T, rG)
        D = red["G"][:, m * (order - 1):, :]
        assert np.count_nonzero(D * (1 - np.eye(m))[None]) == 0
        # (c) the reduction keeps every cell's box
        rad = np.abs(red["G"]).sum(axis=1)
        np.testing.assert_allclose(red["c"] - rad, got["lo"], rtol=0, atol=BOX_ATOL)
        np.testing.assert_allclose(red["c"] + rad, got["hi"], rtol=0, atol=BOX_ATOL)
    # (d) end to end against the CPU oracle pipeline (DeepZ + reduction)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii, keep_cap=256)
    _, refG = _oracle().reduce_order(ref["ncols"], ref["c"], ref["G"], 2)
    red = ctx.fetch_reduced(N, m, 2)
    scale = np.abs(refG).max()
    np.testing.assert_allclose(red["G"], refG, rtol=0, atol=1e-12 * scale)


@pytest.mark.parametrize("name,lo,hi,parts,cap", [
    ("TORA_ReLU", [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6], [3, 3, 2, 2], 24),      # 32-slot sort: one key per lane
    ("TORA_ReLU", [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6], [3, 3, 2, 2], 40),      # two keys per lane
    ("TORA_sigmoid", [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6], [3, 2, 2, 2], 100),  # four keys per lane, 64 generators
    ("Quadrotor", [-0.004] * 6 + [0.0] * 6, [0.004] * 6 + [0.0] * 6, [2, 2, 1, 1, 1, 1] + [1] * 6, 300)])   # 512 slots: the shared-memory sort
def test_order_reduction_every_sort_width(ctx, name, lo, hi, parts, cap):
    """clr_deepz_fetch_reduced picks its kernel by the generator capacity (keys sorted in registers up to 256 slots, in shared
    memory beyond): every width against the oracle's twin on the same zonotopes, bit for bit."""
    import clr_b200
    net = load_net(name)
    g = clr_b200.split_box(lo, hi, parts)
    got = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii, keep_cap=cap)
    N, m = got["lo"].shape
    for order in (1, 2, 3):
        if got["ncols"].min() < m * (order - 1):
            continue
        red = ctx.fetch_reduced(N, m, order)
        oc, oG = _oracle().reduce_order(got["ncols"], got["c"], got["G"], order)
        assert np.array_equal(red["c"], oc) and np.array_equal(red["G"], oG)


def test_order_reduction_errors(ctx):
    net = load_net("Quadrotor")
    dn = ctx.upload(net)
    c = np.zeros((2, 12)); G = 0.01 * np.eye(12)[:2]          # cells with a single generator
    r = ctx.deepz_zonotopes(dn, c, [0, 1, 2], G, keep_cap=256)
    assert r["ncols"].max() >= 3                               # sigmoid layers append generators: order 2 is fine
    ctx.fetch_reduced(2, 3, 2)
    with pytest.raises(Exception):
        ctx.fetch_reduced(2, 3, 200)                           # needs 597 generators: the reference throws here
    with pytest.raises(Exception):
        ctx.fetch_reduced(2, 3, 0)


def test_project_oa_batched(ctx):
    """_project_oa(R::AbstractTaylorModelReachSet, vars, t) (src/setops.jl:9-12) for a batch of reach-sets: against the
    NumPy restatement, the enclosure property at full size, and chained into the controller propagation."""
    from oracle import deepz_np as dz
    from test_oracle import _random_taylor_models
    rng = np.random.default_rng(21)
    N, n, oT, oQ = 64, 6, 8, 3                              # TORA-like: 4 states + disturbance/control variables, TMJets orders
    exps, coeffs, rem, dt = _random_taylor_models(rng, N, n, oT, oQ)
    vk = [0, 1, 2, 3]
    got = ctx.project_oa(exps, coeffs, rem, dt, vk)
    for i in range(N):
        c, G = dz.project_oa_taylor_model(exps, coeffs[i], rem[i], dt[i], vk)
        p = got["ncols"][i]
        assert p == G.shape[1]
        scale = max(np.abs(G).max(), np.abs(c).max())
        np.testing.assert_allclose(got["c"][i], c, rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(got["G"][i, :p].T, G, rtol=0, atol=1e-12 * scale)
    # zero columns kept on request: always 2 n columns
    got0 = ctx.project_oa(exps, coeffs, rem, dt, vk, remove_zero_generators=False)
    assert np.all(got0["ncols"] == 2 * n)
    np.testing.assert_allclose(np.abs(got0["G"]).sum(axis=1), np.abs(got["G"]).sum(axis=1), rtol=0, atol=1e-12)
    # enclosure at a larger size: every sampled point of every Taylor model lies in the box of its zonotope
    N2 = 20000
    exps, coeffs, rem, dt = _random_taylor_models(rng, N2, n, 4, 2)
    big = ctx.project_oa(exps, coeffs, rem, dt, vk)
    rad = np.abs(big["G"]).sum(axis=1)
    lo, hi = big["c"] - rad, big["c"] + rad
    tpow = dt[:, None] ** np.arange(5)[None, :]
    for _ in range(4):
        xi = rng.uniform(-1, 1, size=(N2, n))
        mono = np.prod(np.power(xi[:, None, :], exps[None, :, :]), axis=2)            # (N2, M)
        val = np.einsum("ivkm,ik,im->iv", coeffs, tpow, mono)[:, vk]
        assert np.all(val + rem[:, vk, 0] >= lo - 1e-12) and np.all(val + rem[:, vk, 1] <= hi + 1e-12)
    # the zonotopes feed the controller propagation of the next control period (src/solve.jl:183-195)
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    sub = slice(0, 500)
    nc = big["ncols"][sub]
    col_off = np.concatenate([[0], np.cumsum(nc)]).astype(np.int64)
    Gp = np.concatenate([big["G"][i, :nc[k]] for k, i in enumerate(range(500))])
    out = ctx.deepz_zonotopes(dn, big["c"][sub], col_off, Gp, post=("shift", -10.0))
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), big["c"][sub], col_off, Gp, post=("shift", -10.0))
    _cmp_boxes(out, ref)
    # the same cells in the padded layout (no host-side compaction): bit-identical boxes
    outp = ctx.deepz_zonotopes_padded(dn, big["c"][sub], big["G"][sub], nc, post=("shift", -10.0))
    assert np.array_equal(outp["lo"], out["lo"]) and np.array_equal(outp["hi"], out["hi"])
    assert outp["stats"]["sum_p_in"] == out["stats"]["sum_p_in"]
    # ... and fully on the device: clr_project_oa -> clr_deepz_zonotopes_padded with CLR_FLAG_DEVICE_IO
    import torch
    from clr_b200 import runtime
    dev = torch.device("cuda", 0)
    M5 = 500
    d_coef = torch.from_numpy(coeffs[:M5]).to(dev); d_rem = torch.from_numpy(rem[:M5]).to(dev); d_dt = torch.from_numpy(dt[:M5]).to(dev)
    d_c = torch.empty((M5, 4), dtype=torch.float64, device=dev); d_G = torch.empty((M5, 2 * n, 4), dtype=torch.float64, device=dev)
    d_nc = torch.empty((M5,), dtype=torch.int32, device=dev)
    d_lo = torch.empty((M5, 1), dtype=torch.float64, device=dev); d_hi = torch.empty_like(d_lo)
    ctx.project_oa_dev(n, 4, exps, M5, d_coef, d_rem, d_dt, vk, True, d_c, d_G, d_nc)
    ctx.deepz_zonotopes_padded_dev(dn, 4, M5, d_c, d_G, d_nc, 2 * n, runtime.PrePostArg(None, ("shift", -10.0)), d_lo, d_hi)
    ctx.synchronize()
    assert np.array_equal(d_lo.cpu().numpy(), out["lo"]) and np.array_equal(d_hi.cpu().numpy(), out["hi"])
    with pytest.raises(Exception):
        ctx.project_oa(exps, coeffs, rem, dt, [7])           # variable outside 0..n-1
    with pytest.raises(Exception):
        ctx.deepz_zonotopes_padded(dn, big["c"][:2], big["G"][:2], np.array([99, 0], dtype=np.int32))   # ncols > cap


def test_project_oa_merges_redundant_generators(ctx):
    """CLR_FLAG_MERGE_REDUNDANT: LazySets' remove_redundant_generators on the full zonotope before the projection (what
    src/setops.jl:10 runs by default), against the NumPy restatement -- Taylor models shaped like TMJets output for a controlled
    plant: the input variables enter their own component only, so their linear and remainder generators are parallel."""
    from oracle import deepz_np as dz
    from test_oracle import _random_taylor_models
    rng = np.random.default_rng(33)
    N, n = 300, 6
    exps, coeffs, rem, dt = _random_taylor_models(rng, N, n, 4, 2)
    deg = exps.sum(axis=1)
    lin = [int(np.flatnonzero((deg == 1) & (exps[:, j] == 1))[0]) for j in range(n)]
    for v in (4, 5):                                   # components 4, 5: u' = 0 -> u(t) = c + s xi_v exactly
        coeffs[:, v, :, :] = 0.0
        coeffs[:, v, 0, 0] = rng.normal(size=N)
        coeffs[:, v, 0, lin[v]] = 0.05
        coeffs[:, :4, :, lin[v]] *= (rng.random((N, 1, 1)) > 0.5)      # in half of the sets the states do not depend on xi_v at all
    coeffs[::7, 1, :, :] = 0.0; rem[::7, 1, :] = 0.0                   # a component that is identically zero now and then
    for vk in ([0, 1, 2, 3], [0, 4], list(range(n))):
        got = ctx.project_oa(exps, coeffs, rem, dt, vk, remove_redundant_generators=True)
        merged = 0
        for i in range(N):
            c, G = dz.project_oa_taylor_model(exps, coeffs[i], rem[i], dt[i], vk, remove_redundant=True)
            _, G0 = dz.project_oa_taylor_model(exps, coeffs[i], rem[i], dt[i], vk, remove_redundant=False)
            merged += G.shape[1] < G0.shape[1]
            p = got["ncols"][i]
            assert p == G.shape[1], (vk, i, p, G.shape)
            scale = max(np.abs(G).max() if G.size else 0.0, np.abs(c).max(), 1e-300)
            np.testing.assert_allclose(got["c"][i], c, rtol=0, atol=1e-12 * scale)
            np.testing.assert_allclose(got["G"][i, :p].T, G, rtol=0, atol=1e-12 * scale)          # signed, column order included
            assert not got["G"][i, p:].any()
        assert merged > 0 or vk == [0, 1, 2, 3]
    plain = ctx.project_oa(exps, coeffs, rem, dt, [0, 4])
    assert np.all(plain["ncols"] >= ctx.project_oa(exps, coeffs, rem, dt, [0, 4], remove_redundant_generators=True)["ncols"])


def test_setops_golden_fixture_on_the_device(ctx):
    """clr_project_oa against the committed fixture (tests/golden/setops_oracle.npz)."""
    z = np.load(os.path.join(GOLDEN, "setops_oracle.npz"))
    got = ctx.project_oa(z["tm_exps"], z["tm_coeffs"], z["tm_rem"], z["tm_dt"], [0, 2])
    for i in range(z["tm_dt"].shape[0]):
        G = z[f"proj{i}_G"]
        assert got["ncols"][i] == G.shape[1]
        scale = max(np.abs(G).max(), np.abs(z[f"proj{i}_c"]).max())
        np.testing.assert_allclose(got["c"][i], z[f"proj{i}_c"], rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(got["G"][i, :G.shape[1]].T, G, rtol=0, atol=1e-12 * scale)


def test_cells_too_large_for_shared_memory_take_the_global_store_pass(ctx):
    """Zonotopes with 400 generators entering TORA's 100-neuron layers: after the first relaxations a cell is several
    hundred columns of 100 rows, more than an SM's shared memory holds: the retry and the global-store passes run (with the
    folded -10 shift applied in the output stage) and still agree with the oracle; a run that keeps the zonotopes (the
    shift as a real layer) returns the same boxes."""
    from clr_b200 import workloads
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    rng = np.random.default_rng(5)
    C, off, G = workloads.random_zonotopes(rng, 40, 4, 380, 400, -0.6, 0.6, 0.004)
    ctx.set_profiling(True)
    try:
        got = ctx.deepz_zonotopes(dn, C, off, G, post=("shift", -10.0), hull=True)
        info = ctx.last_pass_info()
    finally:
        ctx.set_profiling(False)
    assert info["big"] > 0, info
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), C, off, G, post=("shift", -10.0))
    _cmp_boxes(got, ref)
    assert got["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"] and got["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"]
    assert got["hull_lo"][0] == got["lo"].min() and got["hull_hi"][0] == got["hi"].max()
    kept = ctx.deepz_zonotopes(dn, C, off, G, post=("shift", -10.0), keep_cap=1024)
    assert np.array_equal(kept["lo"], got["lo"]) and np.array_equal(kept["hi"], got["hi"])


def test_device_side_zonotope_and_sign_splitters(ctx):
    """clr_split_zonotopes / clr_sign_split (src/splitters.jl:30-43, :60-71) against the host mirror (clr_b200/splitters.py, itself
    the restatement of LazySets split(Z, gens, nparts)): children bit for bit and in the same order; chained into the
    controller propagation."""
    import clr_b200
    rng = np.random.default_rng(17)
    N, n, cap = 37, 4, 9
    ncols = rng.integers(5, cap + 1, N).astype(np.int32)
    Cc = rng.uniform(-0.7, 0.7, (N, n))
    G = np.zeros((N, cap, n))
    for i in range(N):
        G[i, :ncols[i]] = 0.05 * rng.uniform(-1, 1, (ncols[i], n))
    gens, nparts = [0, 3, 4, 1], [2, 1, 0, 1]
    got = ctx.split_zonotopes(Cc, G, ncols, gens, nparts)
    H = sum(nparts)
    assert got["c"].shape == (N << H, n)
    for i in range(N):
        ref = clr_b200.split_zonotope(Cc[i], G[i, :ncols[i]].T, gens, nparts)
        assert len(ref) == 1 << H
        for b in range(1 << H):
            c, Gr = ref.cell(b)
            k = (i << H) + b
            assert np.array_equal(got["c"][k], c), (i, b)
            assert got["ncols"][k] == ncols[i]
            assert np.array_equal(got["G"][k, :ncols[i]].T, Gr), (i, b)
            assert not got["G"][k, ncols[i]:].any()
    # the children go straight into the controller propagation (padded layout)
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    out = ctx.deepz_zonotopes_padded(dn, got["c"], got["G"], got["ncols"], post=("shift", -10.0))
    off = np.concatenate([[0], np.cumsum(got["ncols"])]).astype(np.int64)
    Gc = np.concatenate([got["G"][k, :got["ncols"][k]] for k in range(len(off) - 1)])
    ref = _oracle().deepz_zonotopes(_oracle().Net(net.as_tuples()), got["c"], off, Gc, post=("shift", -10.0))
    _cmp_boxes(out, ref)
    with pytest.raises(Exception):
        ctx.split_zonotopes(Cc, G, ncols, [8], [1])              # generator 8 does not exist in the cells with fewer columns
    # SignSplitter on the control intervals of a split
    lo = np.array([-1.0, 0.0, -3.0, -1.0, 0.5, -2.0, -0.0, 1e-300]); hi = np.array([2.0, 2.0, -1.0, 0.0, 0.7, 5.0, 1.0, 2e-300])
    l2, h2, par = ctx.sign_split(lo, hi)
    exp = []
    for i, (a, b) in enumerate(zip(lo, hi)):
        exp += [(x, y, i) for x, y in clr_b200.SignSplitter().apply(float(a), float(b))]
    assert [(float(a), float(b), int(p)) for a, b, p in zip(l2, h2, par)] == exp
    big_lo = rng.uniform(-1, 0.5, 100003); big_hi = big_lo + rng.uniform(0, 1, 100003)
    l3, h3, p3 = ctx.sign_split(big_lo, big_hi)
    straddle = (big_lo < 0) & (big_hi > 0)
    assert len(l3) == 100003 + straddle.sum() and np.all(np.diff(p3) >= 0)
    assert np.array_equal(np.bincount(p3, minlength=100003), 1 + straddle.astype(np.int64))
    first = np.concatenate([[True], np.diff(p3) > 0])
    assert np.array_equal(l3[first], big_lo) and np.all(h3[first][straddle] == 0.0)


def test_narrow_networks_take_the_one_cta_per_cell_kernel(ctx):
    """Networks whose layers are all <= 32 neurons wide (ACC, VerticalCAS, InvertedPendulum: the reference's CPU-runnable sizes) run
    deepz_tiny_kernel when no statistics / zonotopes are requested: against the oracle, against the tiled kernel (to rounding),
    box grids and zonotope batches, pre- and post-processing, hull, and a large batch."""
    import clr_b200
    from clr_b200 import workloads
    rng = np.random.default_rng(41)
    A = np.zeros((5, 6)); A[2, 4] = 1.0; A[3, 0], A[3, 3] = 1.0, -1.0; A[4, 1], A[4, 4] = 1.0, -1.0
    d = np.array([30.0, 1.4, 0.0, 0.0, 0.0])
    cases = [("ACC_relu", (A, d), None, 6), ("ACC_tanh", (A, d), None, 6), ("InvertedPendulum", None, None, 2),
             ("VerticalCAS_3", None, ("matrix", np.array([[1.0, -1.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 2.0]])), None)]
    for name, pre, post, n0 in cases:
        net = load_net(name)
        n = n0 or net.dim_in
        dn = ctx.upload(net)
        on = _oracle().Net(net.as_tuples())
        C, off, G = workloads.random_zonotopes(rng, 257, n, 0, 14, 0.0, 1.0, 0.05)
        if name.startswith("ACC"):
            C += np.array([90.0, 32.0, 0.0, 10.0, 30.0, 0.0])
        l0 = ctx.launch_count
        got = ctx.deepz_zonotopes(dn, C, off, G, pre=pre, post=post, stats=False, hull=True)
        assert ctx.launch_count - l0 == 1, name                      # ONE kernel: no pilot, no retry pass
        ref = _oracle().deepz_zonotopes(on, C, off, G, pre=pre, post=post)
        scale = max(1.0, np.abs(ref["hi"]).max(), np.abs(ref["lo"]).max())
        np.testing.assert_allclose(got["lo"], ref["lo"], rtol=0, atol=BOX_ATOL * scale, err_msg=name)
        np.testing.assert_allclose(got["hi"], ref["hi"], rtol=0, atol=BOX_ATOL * scale, err_msg=name)
        np.testing.assert_array_equal(got["hull_lo"], got["lo"].min(0))
        np.testing.assert_array_equal(got["hull_hi"], got["hi"].max(0))
        ctx.set_tiny(False)
        try:
            tiled = ctx.deepz_zonotopes(dn, C, off, G, pre=pre, post=post, stats=False)
        finally:
            ctx.set_tiny(True)
        np.testing.assert_allclose(got["lo"], tiled["lo"], rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(got["hi"], tiled["hi"], rtol=0, atol=1e-12 * scale)
        # the padded layout and a single zonotope (the C1 case: one cell per control period)
        one = ctx.deepz_zonotopes(dn, C[:1], off[:2], G[:off[1]], pre=pre, post=post, stats=False)
        assert np.array_equal(one["lo"], got["lo"][:1]) and np.array_equal(one["hi"], got["hi"][:1])
    # box grid with a flat dimension and a cut-point partition, 50 000 cells, shards bit-identical
    net = load_net("InvertedPendulum")
    dn = ctx.upload(net)
    g = clr_b200.split_box([1.0, 0.0], [1.2, 0.2], [250, 200])
    full = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, stats=False, hull=True)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g.partition, g.centers, g.radii, cell_begin=31000, cell_count=3000)
    np.testing.assert_allclose(full["lo"][31000:34000], ref["lo"], rtol=0, atol=BOX_ATOL)
    np.testing.assert_allclose(full["hi"][31000:34000], ref["hi"], rtol=0, atol=BOX_ATOL)
    part = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=12345, cell_count=777, stats=False)
    assert np.array_equal(part["lo"], full["lo"][12345:13122]) and np.array_equal(part["hi"], full["hi"][12345:13122])
    assert full["hull_lo"][0] == full["lo"].min()
    g2 = clr_b200.split_box([1.0, 0.1], [1.2, 0.1], [4, 1])
    flat = ctx.deepz_boxgrid(dn, g2.partition, g2.centers, g2.radii, stats=False)
    ref = _oracle().deepz_boxgrid(_oracle().Net(net.as_tuples()), g2.partition, g2.centers, g2.radii)
    _cmp_boxes(flat, ref)
    # statistics or kept zonotopes: the tiled kernel, same boxes to rounding
    st = ctx.deepz_boxgrid(dn, g2.partition, g2.centers, g2.radii, stats=True)
    np.testing.assert_allclose(st["lo"], flat["lo"], rtol=0, atol=1e-13)
