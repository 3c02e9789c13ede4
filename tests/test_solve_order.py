"""Ordering contract of the breadth-first ``_solve`` (clr_b200/solve.py) against a LITERAL re-implementation of the
reference's depth-first walk (/root/reference/src/solve.jl:145-205: waiting list, ``pop!`` at :179, append order
:168-175 / :197-204), with an ``IndexedSplitter`` firing at k > 1 (src/splitters.jl:49-54) and an input splitter applied to
every control box of every period (src/solve.jl:232).  BASELINE north_star: split indexing and cell ordering bit-exact."""
import numpy as np
import pytest

from conftest import load_net

TORA_LO, TORA_HI = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]


# ---- a stand-in for the plant step (post + _project_oa stay on the reference's CPU path) ---------------------------
_A = np.array([[1.0, 0.1, 0.0, 0.0], [-0.1, 1.0, 0.01, 0.0], [0.0, 0.0, 1.0, 0.1], [0.0, 0.0, 0.0, 1.0]])
_B = np.array([0.0, 0.0, 0.005, 0.1])


def _plant_one(k, c, G, ulo, uhi):
    um, ur = (ulo[0] + uhi[0]) / 2, (uhi[0] - ulo[0]) / 2
    c2 = _A @ c + _B * (um + 10.0) * 0.02 + 0.001 * k
    G2 = np.hstack([_A @ G, (_B * ur * 0.02)[:, None], 0.002 * np.eye(4)[:, :2]])
    return c2, G2


def plant_step(k, cell_c, cell_G, U_lo, U_hi):
    out = [_plant_one(k, cell_c[i], cell_G[i], U_lo[i], U_hi[i]) for i in range(len(cell_G))]
    return np.array([o[0] for o in out]), [o[1] for o in out]


def _sign_split(lo, hi):
    if lo < 0.0 and hi > 0.0:
        return [(lo, 0.0), (0.0, hi)]
    return [(lo, hi)]


def depth_first_reference(controller_one, X0_lo, X0_hi, periods, splitter, use_sign, shift_for_sign=0.0):
    """src/solve.jl:145-205, literally: flowpipes appended per `_solve_one`, successors pushed, the LAST popped first."""
    import clr_b200
    flowpipes = []                       # (k, cell_c, cell_G, U_lo, U_hi)
    waiting = []                         # (end_c, end_G, k)

    def cells_of(lo, hi, k):
        if splitter is not None and splitter.haskey(k):
            g = splitter[k].apply(lo, hi)
        else:
            g = clr_b200.split_box(lo, hi, [1] * len(lo))
        out = []
        for i in range(len(g)):
            c, r = g.cell(i)
            out.append((c, np.diag(r)[:, r != 0.0]))
        return out

    def solve_one(k, c, G):
        lo, hi = controller_one(c, G)
        pieces = _sign_split(float(lo[0]), float(hi[0])) if use_sign else [(float(lo[0]), float(hi[0]))]
        return [(k, c, G, np.array([a]), np.array([b])) for a, b in pieces]

    def collect(results, k):
        for Fs in results:
            flowpipes.extend(Fs)
            if k < periods:
                for F in Fs:
                    end_c, end_G = _plant_one(k, F[1], F[2], F[3], F[4])
                    waiting.append((end_c, end_G, k))

    k = 1
    X0s = cells_of(np.asarray(X0_lo, float), np.asarray(X0_hi, float), 1)
    collect([solve_one(1, c, G) for c, G in X0s], 1)
    while waiting:
        end_c, end_G, kp = waiting.pop()
        k = kp + 1
        if splitter is not None and splitter.haskey(k):
            r = np.abs(end_G).sum(axis=1)
            X0s = cells_of(end_c - r, end_c + r, k)
        else:
            X0s = [(end_c, end_G)]
        collect([solve_one(k, c, G) for c, G in X0s], k)
    return flowpipes


def _compare(seq, ref, atol):
    assert len(seq) == len(ref)
    for i, (a, b) in enumerate(zip(seq, ref)):
        assert a[0] == b[0], (i, "control period differs")
        # the cells of k > 1 are functions of the upstream control boxes (which the two walks compute with different, equally
        # valid summation orders): same cell <=> equal to 1e-12; neighbouring cells differ by ~1e-2
        assert a[2].shape == b[2].shape, (i, "cell differs: ordering broken")
        assert np.abs(a[1] - b[1]).max() <= 1e-12 and (a[2].size == 0 or np.abs(a[2] - b[2]).max() <= 1e-12), \
            (i, "cell differs: ordering broken")
        np.testing.assert_allclose(a[3], b[3], rtol=0, atol=atol)
        np.testing.assert_allclose(a[4], b[4], rtol=0, atol=atol)


def test_lifo_emission_order_on_random_trees():
    from clr_b200.solve import lifo_emission_order
    rng = np.random.default_rng(3)
    for trial in range(50):
        depth = int(rng.integers(1, 6))
        sizes = [int(rng.integers(1, 5))]
        parents = [sizes[0]]
        for k in range(1, depth):
            counts = rng.integers(0, 4, sizes[-1])           # a chain may die out (no successor cells)
            par = np.repeat(np.arange(sizes[-1]), counts)
            parents.append(par)
            sizes.append(len(par))
        order = lifo_emission_order(parents)
        # literal waiting-list walk over the same tree
        ref = [(0, i) for i in range(sizes[0])]
        stack = list(ref) if depth > 1 else []
        while stack:
            k, i = stack.pop()
            ch = [j for j in range(sizes[k + 1]) if parents[k + 1][j] == i]
            ref.extend((k + 1, j) for j in ch)
            if k + 1 < depth - 1:
                stack.extend((k + 1, j) for j in ch)
        assert order == ref
        assert sorted(order) == [(k, i) for k in range(depth) for i in range(sizes[k])]      # every flowpipe exactly once


def test_lifo_emission_order_small_example():
    from clr_b200.solve import lifo_emission_order
    # 2 cells at k = 1; each has 2 successors at k = 2; those have 1 successor each at k = 3
    order = lifo_emission_order([2, np.array([0, 0, 1, 1]), np.array([0, 1, 2, 3])])
    # k=1: both; pop cell 1 -> its successors (2, 3); pop 3 -> its successor; pop 2 -> ...; then cell 0
    assert order == [(0, 0), (0, 1), (1, 2), (1, 3), (2, 3), (2, 2), (1, 0), (1, 1), (2, 1), (2, 0)]
    with pytest.raises(ValueError):
        lifo_emission_order([2, np.array([1, 0])])


def _oracle_controllers(net, post):
    from oracle import cbind, deepz_np as dz
    from clr_b200.splitters import BoxGrid
    on = cbind.Net(net.as_tuples())

    def batch(cells):
        if isinstance(cells, BoxGrid):
            r = cbind.deepz_boxgrid(on, cells.partition, cells.centers, cells.radii, post=post)
        else:
            r = cbind.deepz_zonotopes(on, cells.C, cells.col_off, cells.G, post=post)
        return r["lo"], r["hi"]

    def one(c, G):
        return dz.controller_forward(net.as_tuples(), c, G, post=post)[2:]
    return batch, one


@pytest.mark.parametrize("use_sign", [False, True])
def test_bfs_solve_emits_the_depth_first_order_cpu(use_sign):
    """Breadth-first walk (one batched controller call per period, here the C oracle) == depth-first reference walk (one
    NumPy-oracle controller_forward per cell), flowpipe by flowpipe, with splitters at k = 1 and k = 3."""
    import clr_b200
    from clr_b200.solve import emitted_sequence, solve_controller_bfs
    net = load_net("TORA_ReLU")
    post = ("shift", -10.0)          # models/TORA/TORA.jl:75: with it the control boxes of X0's cells straddle 0
    batch, one = _oracle_controllers(net, post)
    splitter = clr_b200.IndexedSplitter({1: clr_b200.BoxSplitter([2, 2, 1, 1]), 3: clr_b200.BoxSplitter([1, 1, 2, 1])})
    levels, order = solve_controller_bfs(batch, TORA_LO, TORA_HI, 4, splitter, plant_step,
                                         input_splitter=clr_b200.SignSplitter() if use_sign else None)
    ref = depth_first_reference(one, TORA_LO, TORA_HI, 4, splitter, use_sign)
    seq = emitted_sequence(levels, order)
    _compare(seq, ref, 1e-12)
    assert len(seq) >= 4 + 4 + 8 + 8
    if use_sign:
        assert len(seq) > 24, "the input splitter never fired: the test would not exercise src/solve.jl:232"
    # a plain splitter fires at k = 1 only (src/splitters.jl:12-13)
    levels, order = solve_controller_bfs(batch, TORA_LO, TORA_HI, 3, clr_b200.BoxSplitter([3, 1, 2, 1]), plant_step)
    ref = depth_first_reference(one, TORA_LO, TORA_HI, 3, clr_b200.BoxSplitter([3, 1, 2, 1]), False)
    _compare(emitted_sequence(levels, order), ref, 1e-12)
    assert [lv.cell_c.shape[0] for lv in levels] == [6, 6, 6]


@pytest.mark.gpu
@pytest.mark.parametrize("use_sign", [False, True])
def test_bfs_solve_emits_the_depth_first_order_gpu(use_sign):
    """The same contract with the CUDA path as the batched controller: one clr_deepz_* call per control period."""
    import clr_b200
    from clr_b200 import runtime
    from clr_b200.solve import emitted_sequence, solve_controller_bfs
    from clr_b200.splitters import BoxGrid
    net = load_net("TORA_ReLU")
    post = ("shift", -10.0)
    _, one = _oracle_controllers(net, post)
    ctx = runtime.Context(0)
    try:
        dn = ctx.upload(net)
        calls = []

        def batch(cells):
            l0 = ctx.launch_count
            if isinstance(cells, BoxGrid):
                r = ctx.deepz_boxgrid(dn, cells.partition, cells.centers, cells.radii, post=post, stats=False)
            else:
                r = ctx.deepz_zonotopes(dn, cells.C, cells.col_off, cells.G, post=post, stats=False)
            calls.append(ctx.launch_count - l0)
            return r["lo"], r["hi"]

        splitter = clr_b200.IndexedSplitter({1: clr_b200.BoxSplitter([4, 4, 3, 5]), 2: clr_b200.BoxSplitter([1, 1, 2, 1]),
                                             4: clr_b200.BoxSplitter([2, 1, 1, 1])})
        K = 5
        levels, order = solve_controller_bfs(batch, TORA_LO, TORA_HI, K, splitter, plant_step,
                                             input_splitter=clr_b200.SignSplitter() if use_sign else None)
        assert len(calls) == K                               # ONE batched call per control period
        ref = depth_first_reference(one, TORA_LO, TORA_HI, K, splitter, use_sign)
        _compare(emitted_sequence(levels, order), ref, 1e-10)
        assert len(ref) >= 240 + 480 + 480 + 960 + 960
    finally:
        ctx.close()
