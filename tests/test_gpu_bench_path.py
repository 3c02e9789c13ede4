"""Parity of the code path bench.py measures: calls of at least CLR_REORDER_MIN_CELLS (16 384) cells take the pilot-ordered
network (reordered neurons, skipped k-steps), the no-statistics lean kernel, the retry pass and -- in the dense regime --
the global-store pass.  Every regime is checked against the CPU oracle on lo AND hi, the reordered against the
unreordered network at 10^6 cells, and the 64-bit cell decode against the oracle's exact integer arithmetic.
Tolerance: output boxes 1e-10 absolute, scaled by the magnitude of the bounds where that exceeds 1 (BASELINE north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOX_ATOL = 1e-10
REORDER_MIN_CELLS = 16384          # clr_api.cu: CLR_REORDER_MIN_CELLS


@pytest.fixture(scope="module")
def ctx():
    from clr_b200 import runtime
    c = runtime.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def c4():
    from clr_b200 import workloads
    from oracle import cbind
    net = workloads.synthetic_relu_net()
    return net, cbind.Net(net.as_tuples())


def _oracle_cells(on, g, ranges):
    from oracle import cbind
    lo, hi = [], []
    for b, cnt in ranges:
        r = cbind.deepz_boxgrid(on, g.partition, g.centers, g.radii, cell_begin=b, cell_count=cnt)
        lo.append(r["lo"]); hi.append(r["hi"])
    return np.concatenate(lo), np.concatenate(hi)


@pytest.mark.parametrize("regime,count,begin", [("stable", 40000, 123456), ("mixed", 20000, 500000), ("dense", 16384, 31000)])
def test_reordered_skipping_path_against_the_oracle(ctx, c4, regime, count, begin):
    """>= 16 384 cells of the bench split in every regime, no statistics (the benchmarked instantiation), lo and hi against
    the oracle for EVERY cell; the passes the regime is known to need must actually have run."""
    from clr_b200 import workloads
    net, on = c4
    assert count >= REORDER_MIN_CELLS
    dn = ctx.upload(net)
    g = workloads.c4_grid(regime)
    ctx.set_profiling(True)
    try:
        got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=begin, cell_count=count, stats=False, hull=True)
        info = ctx.last_pass_info()
    finally:
        ctx.set_profiling(False)
    if regime == "dense":
        assert info["big"] > 0, info        # cells wider than an SM's shared memory: the global-store pass ran
    if regime != "stable":
        assert info["retried"] + info["big"] > 0, info
    lo, hi = _oracle_cells(on, g, [(begin, count)])
    scale = max(1.0, np.abs(hi).max(), np.abs(lo).max())
    np.testing.assert_allclose(got["lo"], lo, rtol=0, atol=BOX_ATOL * scale)
    np.testing.assert_allclose(got["hi"], hi, rtol=0, atol=BOX_ATOL * scale)
    assert got["hull_lo"][0] == got["lo"].min() and got["hull_hi"][0] == got["hi"].max()
    # the statistics instantiation on the same (reordered) path: the reference's generator counts, cell by cell summed
    st = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=begin, cell_count=count, stats=True)
    assert np.array_equal(st["lo"], got["lo"]) and np.array_equal(st["hi"], got["hi"])
    from oracle import cbind
    ref = cbind.deepz_boxgrid(on, g.partition, g.centers, g.radii, cell_begin=begin, cell_count=count)
    assert st["stats"]["sum_p_in"] == ref["stats"]["sum_p_in"]
    assert st["stats"]["sum_unstable"] == ref["stats"]["sum_unstable"]


@pytest.mark.parametrize("regime", ["stable", "mixed", "dense"])
def test_reorder_on_equals_reorder_off_at_full_size(ctx, c4, regime):
    """10^6 cells (the bench size): the pilot-ordered network against the network as uploaded.  The reordering only changes
    the summation order inside the layer products: <= 1e-12 relative to the magnitude of the bounds, and both runs agree
    with the oracle on cells sampled over the whole split."""
    from clr_b200 import workloads
    net, on = c4
    dn = ctx.upload(net)
    g = workloads.c4_grid(regime)
    assert len(g) == 10 ** 6
    ctx.set_reorder(True)
    a = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, stats=False)
    ctx.set_reorder(False)
    try:
        b = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, stats=False)
    finally:
        ctx.set_reorder(True)
    scale = max(1.0, np.abs(b["hi"]).max(), np.abs(b["lo"]).max())
    assert np.abs(a["lo"] - b["lo"]).max() <= 1e-12 * scale
    assert np.abs(a["hi"] - b["hi"]).max() <= 1e-12 * scale
    assert np.all(a["lo"] <= a["hi"])
    per = {"stable": 400, "mixed": 200, "dense": 40}[regime]
    ranges = [(s, per) for s in range(0, 10 ** 6 - per, 10 ** 6 // 16)]
    lo, hi = _oracle_cells(on, g, ranges)
    idx = np.concatenate([np.arange(s, s + per) for s, _ in ranges])
    np.testing.assert_allclose(a["lo"][idx], lo, rtol=0, atol=BOX_ATOL * scale)
    np.testing.assert_allclose(a["hi"][idx], hi, rtol=0, atol=BOX_ATOL * scale)
    np.testing.assert_allclose(b["lo"][idx], lo, rtol=0, atol=BOX_ATOL * scale)
    np.testing.assert_allclose(b["hi"][idx], hi, rtol=0, atol=BOX_ATOL * scale)


def test_splits_with_more_than_2_pow_32_cells(ctx):
    """A split whose total exceeds 2^32 cells: the strides of the high dimensions do not fit 32 bits even where the cell id
    does (the 32-bit fast path of the cell decode truncated them).  Small ranges near id 0, across the 2^32
    boundary and at the far end, against the oracle (exact integer decode)."""
    import clr_b200
    from conftest import load_net
    from oracle import cbind
    net = load_net("TORA_ReLU")
    dn = ctx.upload(net)
    on = cbind.Net(net.as_tuples())
    part = [70000, 70000, 3, 2]                       # 2.94e10 cells; stride[2] = 4.9e9 > 2^32
    g = clr_b200.split_box([0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6], part)
    total = len(g)
    assert total > 2 ** 32
    for begin in (0, 69990, 2 ** 32 - 40, 70000 * 70000 - 20, 70000 * 70000 * 3 - 15, total - 64):
        cnt = 64
        got = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=begin, cell_count=cnt, post=("shift", -10.0))
        ref = cbind.deepz_boxgrid(on, g.partition, g.centers, g.radii, cell_begin=begin, cell_count=cnt, post=("shift", -10.0))
        np.testing.assert_allclose(got["lo"], ref["lo"], rtol=0, atol=BOX_ATOL, err_msg=str(begin))
        np.testing.assert_allclose(got["hi"], ref["hi"], rtol=0, atol=BOX_ATOL, err_msg=str(begin))
        # and against the Python-int decode of the host mirror: the cell's own box must contain its centre's control
        c, r = g.cell(begin + 5)
        u = ctx.forward_points(dn, c[None, :], post=("shift", -10.0))[0, 0]
        assert got["lo"][5, 0] - 1e-12 <= u <= got["hi"][5, 0] + 1e-12


def test_block_cyclic_shards_reassemble_to_the_contiguous_result(ctx, c4):
    """clr_deepz_boxgrid_cyclic: rank r of G takes the blocks r, r + G, ... of `block` cells; the shards put back in id order
    equal one contiguous call bit for bit (same pilot order for every shard: it is derived from the whole split)."""
    from clr_b200 import workloads
    net, on = c4
    dn = ctx.upload(net)
    g = workloads.c4_grid("mixed")
    first, N, G, blk = 250000, 70000, 3, 1024
    full = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=first, cell_count=N, stats=False)
    lo = np.full((N, 1), np.nan)
    hi = np.full((N, 1), np.nan)
    nb = (N + blk - 1) // blk
    for r in range(G):
        ids = np.concatenate([np.arange(j * blk, min(N, (j + 1) * blk)) for j in range(r, nb, G)])
        part = ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=first + r * blk, cell_count=len(ids), stats=False,
                                 block=blk, stride=G * blk, hull=True)
        lo[ids], hi[ids] = part["lo"], part["hi"]
        assert part["hull_lo"][0] == part["lo"].min()
    assert np.array_equal(lo, full["lo"]) and np.array_equal(hi, full["hi"])
    from clr_b200 import runtime
    with pytest.raises(runtime.ClrArgumentError):      # the last block would lie beyond the split
        ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=10 ** 6 - 2000, cell_count=2000, block=1024, stride=4096)
    with pytest.raises(runtime.ClrArgumentError):
        ctx.deepz_boxgrid(dn, g.partition, g.centers, g.radii, cell_begin=0, cell_count=10, block=8, stride=4)
