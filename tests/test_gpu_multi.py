"""clr_create_multi: several devices behind one handle, world size 2 WITHOUT torch / NCCL.  The GPU box of
the test tier has one B200, so the two shards run as two contexts (two host threads, two streams) on device 0; on a box with
more devices CLR_TEST_DEVICES=0,1,... uses them."""
import os

import numpy as np
import pytest

from conftest import load_net

pytestmark = pytest.mark.gpu
TORA_LO, TORA_HI = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]


def _devices():
    env = os.environ.get("CLR_TEST_DEVICES")
    return [int(x) for x in env.split(",")] if env else [0, 0]


@pytest.fixture(scope="module")
def mctx():
    from clr_b200 import runtime
    m = runtime.MultiContext(_devices())
    yield m
    m.close()


def test_sharded_boxgrid_equals_single_device(mctx):
    import clr_b200
    from clr_b200 import runtime
    net = load_net("TORA_ReLU")
    g = clr_b200.split_box(TORA_LO, TORA_HI, [10, 10, 10, 3])
    ctx = runtime.Context(_devices()[0])
    try:
        one = ctx.deepz_boxgrid(ctx.upload(net), g.partition, g.centers, g.radii, post=("shift", -10.0), hull=True)
    finally:
        ctx.close()
    mn = mctx.upload(net)
    two = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, post=("shift", -10.0), hull=True)
    assert np.array_equal(two["lo"], one["lo"]) and np.array_equal(two["hi"], one["hi"])       # cell order bit-exact
    assert two["hull_lo"][0] == one["hull_lo"][0] and two["hull_hi"][0] == one["hull_hi"][0]    # gathered hull = union
    assert two["stats"]["sum_p_in"] == one["stats"]["sum_p_in"] and two["stats"]["flops"] == one["stats"]["flops"]
    G = len(mctx)
    assert list(two["shard_bounds"]) == [(3000 * r) // G for r in range(G + 1)]
    assert all(c > 0 for c in mctx.launch_count())                                             # every device did launch
    # a sub-range, and an empty one
    part = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, cell_begin=701, cell_count=333, post=("shift", -10.0), hull=True)
    assert np.array_equal(part["lo"], one["lo"][701:1034])
    assert part["hull_lo"][0] == one["lo"][701:1034].min()
    empty = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, cell_begin=5, cell_count=0, post=("shift", -10.0), hull=True)
    assert empty["lo"].shape == (0, 1) and empty["hull_lo"][0] == np.inf


def test_cost_balanced_shards_cover_the_range_in_order(mctx):
    from clr_b200 import workloads
    from oracle import cbind
    net = workloads.synthetic_relu_net()
    g = workloads.c4_grid("mixed")
    mn = mctx.upload(net)
    N = 60000
    mctx.set_reorder(False)
    try:
        bal = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, cell_begin=400000, cell_count=N, hull=True, balance=True, stats=False)
        eq = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, cell_begin=400000, cell_count=N, hull=True, balance=False, stats=False)
        cyc = mctx.deepz_boxgrid(mn, g.partition, g.centers, g.radii, cell_begin=400000, cell_count=N, hull=True, balance="cyclic", stats=True)
    finally:
        mctx.set_reorder(True)
    b = bal["shard_bounds"]
    assert b[0] == 400000 and b[-1] == 400000 + N and np.all(np.diff(b) >= 0)
    assert np.array_equal(bal["lo"], eq["lo"]) and np.array_equal(bal["hi"], eq["hi"])        # boundaries do not change results
    assert np.array_equal(cyc["lo"], eq["lo"]) and np.array_equal(cyc["hi"], eq["hi"])        # block-cyclic shards: same cells, same order
    assert cyc["hull_lo"][0] == eq["hull_lo"][0] and cyc["stats"]["n_cells"] == N
    assert bal["hull_hi"][0] == bal["hi"].max()
    ref = cbind.deepz_boxgrid(cbind.Net(net.as_tuples()), g.partition, g.centers, g.radii, cell_begin=400000 + N - 500, cell_count=500)
    np.testing.assert_allclose(bal["lo"][-500:], ref["lo"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(bal["hi"][-500:], ref["hi"], rtol=0, atol=1e-10)


def test_sharded_rollout_equals_single_device(mctx):
    from clr_b200 import runtime, workloads
    net = load_net("Unicycle")
    rng = np.random.default_rng(2)
    x0, Wd = workloads.unicycle_samples(rng, 1001, 6)
    ctx = runtime.Context(_devices()[0])
    try:
        one = ctx.rollout(ctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 1.2, 6, post=("shift", -20.0))
    finally:
        ctx.close()
    two = mctx.rollout(mctx.upload(net), "Unicycle", x0, Wd, 0.0, 0.2, 1.2, 6, post=("shift", -20.0))
    assert np.array_equal(two["X_end"], one["X_end"]) and np.array_equal(two["U"], one["U"])
    assert np.array_equal(two["naccept"], one["naccept"])


def test_multi_errors(mctx):
    from clr_b200 import runtime
    with pytest.raises(runtime.ClrError):
        runtime.MultiContext([0, 99])
    import clr_b200
    net = load_net("TORA_ReLU")
    g3 = clr_b200.split_box([0, 0, 0], [1, 1, 1], [2, 2, 2])
    with pytest.raises(runtime.ClrError):
        mctx.deepz_boxgrid(mctx.upload(net), g3.partition, g3.centers, g3.radii)          # 3-dim cells into a 4-input network
