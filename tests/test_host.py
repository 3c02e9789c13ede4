"""CPU tests of the host side: the C-ABI library loads and exports what the header declares, the
splitter mirror, the POLAR reader, and the multi-GPU sharding logic over gloo (world size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_net


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "clr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from clr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 18
    for s in declared:
        assert hasattr(L, s), f"{s} is declared in include/clr_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (no oracle / NumPy fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from clr_b200 import runtime
    with pytest.raises(runtime.ClrError) as e:
        runtime.Context(0)
    assert e.value.code == runtime.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "closedloopreachability.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test oracle", "").replace("CPU oracle", "").lower() or f == "network.py", f


def test_box_splitter_matches_oracle_tables_and_order():
    import clr_b200
    from oracle import deepz_np as dz
    lo, hi = [0.6, -0.7, -0.4, 0.5], [0.7, -0.6, -0.3, 0.6]
    for part in ([4, 4, 3, 5], [10, 10, 10, 1], [1, 1, 1, 1], [7, 3, 1, 2]):
        g = clr_b200.BoxSplitter(part).apply(lo, hi)
        cent, rad = dz.split_tables(lo, hi, part)
        for d in range(4):
            assert np.array_equal(g.centers[d], cent[d])          # bit-exact tables
            assert np.all(g.radii[d] == rad[d])
        cells, _ = dz.split_box(lo, hi, part)
        for i in sorted({0, min(1, len(g) - 1), len(g) // 2, len(g) - 1}):
            c, r = g.cell(i)
            assert np.array_equal(c, cells[i])                    # product order, dimension 1 fastest
            assert g.multi_index(i) == dz.cell_multi_index(i, part)
    assert len(clr_b200.BoxSplitter().apply(lo, hi)) == 16            # default: 2 blocks per dimension
    assert clr_b200.julia_range(0.605, 0.01, 10)[2] == 0.625
    sp = clr_b200.BoxSplitter([2, 2, 2, 2])
    assert sp.haskey(1) and not sp.haskey(2)                           # src/splitters.jl:12
    with pytest.raises(KeyError):
        sp[2]
    with pytest.raises(ValueError):
        clr_b200.split_box([0], [1], [0])
    idx = clr_b200.IndexedSplitter({1: sp, 3: clr_b200.NoSplitter()})
    assert idx.haskey(3) and not idx.haskey(2) and len(idx[3].apply(lo, hi)) == 1


def test_cut_point_split():
    import clr_b200
    g = clr_b200.split_box([1.0, 0.0], [1.2, 0.2], [[1.1, 1.16], [0.09, 0.145, 0.18]])
    assert list(g.partition) == [3, 4]
    np.testing.assert_allclose(g.centers[0], [1.05, 1.13, 1.18])
    np.testing.assert_allclose(g.radii[0], [0.05, 0.03, 0.02])
    c, r = g.cell(4)          # second block of dim 2, second block of dim 1
    np.testing.assert_allclose(c, [1.13, 0.1175])


def test_read_polar_roundtrip(tmp_path):
    import clr_b200
    net = load_net("InvertedPendulum")
    names = {"ReLU": "ReLU", "Id": "Affine", "Sigmoid": "sigmoid", "Tanh": "tanh"}
    lines = [str(net.dim_in), str(net.dim_out), str(len(net.layers) - 1)]
    lines += [str(l.dim_out) for l in net.layers[:-1]]
    lines += [names[l.activation] for l in net.layers]
    for l in net.layers:
        for i in range(l.dim_out):
            lines += [repr(float(x)) for x in l.weights[i]] + [repr(float(l.bias[i]))]
    lines += ["0.0", "1.0"]
    p = tmp_path / "net.polar"
    p.write_text("\n".join(lines) + "\n")
    back = clr_b200.read_POLAR(str(p))
    assert back.dims == net.dims
    for a, b in zip(back.layers, net.layers):
        assert np.array_equal(a.weights, b.weights) and np.array_equal(a.bias, b.bias) and a.activation == b.activation
    p.write_text("\n".join(lines[:-3]) + "\n")
    with pytest.raises(ValueError):
        clr_b200.read_POLAR(str(p))


def test_shard_ranges_cover_everything():
    from clr_b200 import parallel
    for n in (0, 1, 7, 1000, 10 ** 6 + 3):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            for (b0, c0), (b1, _) in zip(r[:-1], r[1:]):
                assert b0 + c0 == b1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CLR_ROOT"])
import clr_b200
from clr_b200 import parallel, workloads
from oracle import cbind      # the checker stands in for the GPU kernels in this CPU-only test

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
net = clr_b200.FeedforwardNetwork.load_npz(os.path.join(os.environ["CLR_ROOT"], "tests", "golden", "nets", "TORA_ReLU.npz"))
g = clr_b200.split_box(*workloads.TORA_X0, [4, 4, 3, 5])
b, cnt = parallel.shard_range(len(g), rank, world)
r = cbind.deepz_boxgrid(cbind.Net(net.as_tuples()), g.partition, g.centers, g.radii, cell_begin=b, cell_count=cnt,
                        post=("shift", -10.0))
lo, hi = parallel.gather_hull(torch.from_numpy(r["lo"].min(0)), torch.from_numpy(r["hi"].max(0)))
tmax = parallel.max_over_ranks(1.0 + rank)
tot = parallel.sum_over_ranks(cnt)
per_rank = parallel.gather_scalars(10.0 * (rank + 1))             # per-rank times of the bench line
assert per_rank == [10.0 * (k + 1) for k in range(world)]
# block-cyclic shards (bench default at N > 1): rank r takes the blocks r, r + world, ... of `blk` cells; the oracle has no
# strided call, so the shard is evaluated block by block -- reassembled it must equal the contiguous result, order included
blk = 16
nb = (len(g) + blk - 1) // blk
mine = []
for j in range(rank, nb, world):
    rr = cbind.deepz_boxgrid(cbind.Net(net.as_tuples()), g.partition, g.centers, g.radii, cell_begin=j * blk,
                             cell_count=min(blk, len(g) - j * blk), post=("shift", -10.0))
    mine.append((j, rr["lo"]))
cyc = [None] * world
dist.all_gather_object(cyc, mine)
parts = [None] * world
dist.all_gather_object(parts, r["lo"])
if rank == 0:
    blocks = sorted([b for part in cyc for b in part], key=lambda t: t[0])
    cyc_lo = np.concatenate([b[1] for b in blocks])
    full = cbind.deepz_boxgrid(cbind.Net(net.as_tuples()), g.partition, g.centers, g.radii, post=("shift", -10.0))
    assert np.array_equal(parallel.concat_shards(parts), full["lo"])
    assert float(lo[0]) == full["lo"].min() and float(hi[0]) == full["hi"].max()
    assert tmax == float(world) and tot == len(g)
    assert np.array_equal(cyc_lo, full["lo"])
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, CLR_ROOT=ROOT, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_zonotope_and_sign_splitters():
    import clr_b200
    from oracle import deepz_np as dz
    rng = np.random.default_rng(4)
    c, G = rng.uniform(-1, 1, 4), rng.uniform(-0.3, 0.3, (4, 5))
    cells = clr_b200.split_zonotope(c, G, [0, 3], [2, 1])
    ref = dz.split_zonotope(c, G, [0, 3], [2, 1])
    assert len(cells) == len(ref) == 8
    for i, (rc, rG) in enumerate(ref):
        cc, GG = cells.cell(i)
        assert np.array_equal(cc, rc) and np.array_equal(GG, rG)          # same cells, same order
    # the cells cover the original zonotope: every vertex combination of the split generators is reproduced
    default = clr_b200.ZonotopeSplitter().apply(c, G)                     # one split per generator
    assert len(default) == 2 ** 5
    lo = (default.C - np.abs(default.G.reshape(32, 5, 4)).sum(1)).min(0)
    np.testing.assert_allclose(lo, c - np.abs(G).sum(1), rtol=1e-14)
    with pytest.raises(ValueError):
        clr_b200.split_zonotope(c, G, [7], [1])
    s = clr_b200.SignSplitter()
    assert s.apply(-1.0, 2.0) == [(-1.0, 0.0), (0.0, 2.0)] and s.apply(0.0, 2.0) == [(0.0, 2.0)]


def test_controlled_plant_and_processing_mirror():
    """apply() semantics of src/problem.jl:5-77 on vectors and the dimension check of src/solve.jl:130-134."""
    import clr_b200
    u = np.array([1.0, -2.0, 3.0])
    assert np.array_equal(clr_b200.NoPostprocessing().apply(u), u)
    assert np.array_equal(clr_b200.UniformAdditivePostprocessing(-10.0).apply(u), u - 10.0)
    assert np.array_equal(clr_b200.ProjectionPostprocessing([3, 1]).apply(u), [3.0, 1.0])      # 1-based like Julia
    assert np.array_equal(clr_b200.LinearMapPostprocessing(11.0).apply(u), 11.0 * u)
    M = np.array([[1.0, 0.0, 2.0]])
    assert np.array_equal(clr_b200.LinearMapPostprocessing(M).apply(u), M @ u)
    assert clr_b200.ProjectionPostprocessing([3, 1]).as_arg() == ("project", [2, 0])
    with pytest.warns(UserWarning):
        clr_b200.UniformAdditivePostprocessing(0.0)                                           # src/problem.jl:16-19
    pre = clr_b200.AffinePreprocessing(np.eye(2), [1.0, 2.0])
    assert np.array_equal(pre.apply([1.0, 1.0]), [2.0, 3.0])
    net = load_net("Unicycle")
    vars_ = {"states": [1, 2, 3, 4], "disturbances": [5], "controls": [6, 7]}
    lo = [9.5, -4.5, 2.1, 1.5, -1e-4, 0.0, 0.0]
    hi = [9.55, -4.45, 2.11, 1.51, 1e-4, 0.0, 0.0]
    prob = clr_b200.ControlledPlant("Unicycle", lo, hi, net, vars_, 0.2,
                                    postprocessing=clr_b200.UniformAdditivePostprocessing(-20.0))
    assert prob.states() == [1, 2, 3, 4] and prob.disturbances() == [5] and prob.controls() == [6, 7]
    assert np.array_equal(prob.project(prob.disturbances())[0], [-1e-4])
    with pytest.raises(ValueError):
        clr_b200.ControlledPlant("Unicycle", lo[:6], hi[:6], net, vars_, 0.2)
    x = clr_b200.sample_box(np.random.default_rng(0), lo[:4], hi[:4], 5, include_vertices=True)
    assert x.shape == (5 + 16, 4) and np.all(x >= np.array(lo[:4])) and np.all(x <= np.array(hi[:4]))


def test_runtime_plant_source_compiles_without_a_device():
    """clr_plant_compile with ctx == NULL: NVRTC syntax check of a user right-hand side (no GPU needed)."""
    from clr_b200 import runtime
    ok = "dx[0] = x[1]; dx[1] = 2.0 * sin(x[0]) + 8.0 * q[0];"          # models/InvertedPendulum/InvertedPendulum.jl:44-51
    runtime.CompiledPlant(None, 2, 0, 1, ok)
    with pytest.raises(runtime.ClrArgumentError) as e:
        runtime.CompiledPlant(None, 2, 0, 1, "dx[0] = no_such_symbol;")
    assert "no_such_symbol" in str(e.value)
    with pytest.raises(runtime.ClrArgumentError):
        runtime.CompiledPlant(None, 0, 0, 1, ok)


def test_cost_weighted_bounds():
    """parallel.bounds_from_costs: contiguous shards of equal estimated cost."""
    from clr_b200 import parallel
    edges = [0, 100, 200, 300, 400]
    assert parallel.bounds_from_costs(edges, [1, 1, 1, 1], 4) == [0, 100, 200, 300, 400]
    assert parallel.bounds_from_costs(edges, [1, 1, 1, 1], 2) == [0, 200, 400]
    b = parallel.bounds_from_costs(edges, [3, 1, 1, 1], 2)            # total 600: half = 300 = the whole first block
    assert b == [0, 100, 400]
    b = parallel.bounds_from_costs(edges, [1, 1, 1, 3], 3)            # total 600, thirds at 200 and 400 cost units
    assert b == [0, 200, 333, 400]
    b = parallel.bounds_from_costs([10, 20], [0.0], 3)                # no cost information: equal counts
    assert b == [10, 13, 16, 20]
    rng = np.random.default_rng(0)
    for _ in range(20):
        B, W = int(rng.integers(1, 40)), int(rng.integers(1, 9))
        e = np.concatenate([[0], np.cumsum(rng.integers(1, 1000, B))]).tolist()
        c = rng.uniform(0.5, 2.0, B).tolist()
        b = parallel.bounds_from_costs(e, c, W)
        assert len(b) == W + 1 and b[0] == e[0] and b[-1] == e[-1] and all(x <= y for x, y in zip(b, b[1:]))
        # shard costs within one block's worth of each other
        cum = lambda x: sum(ci * (min(x, e[i + 1]) - e[i]) for i, ci in enumerate(c) if x > e[i])
        costs = [cum(b[r + 1]) - cum(b[r]) for r in range(W)]
        assert max(costs) - min(costs) <= 2.0 * 1000 + 1e-9


def test_julia_shim_wraps_every_exported_symbol():
    """julia/ClrB200.jl holds a `ccall` for every symbol include/clr_b200.h declares; the Julia files
    cannot be executed in this image, so at least the binding surface is checked for completeness."""
    import re
    src = open(os.path.join(ROOT, "julia", "ClrB200.jl")).read()
    bound = set(re.findall(r"\(:(clr_[a-z0-9_]+), LIB\)", src))
    missing = sorted(set(_header_symbols()) - bound)
    assert not missing, f"no Julia wrapper for {missing}"
    for f in ("solve_b200.jl", "simulate_b200.jl"):
        assert os.path.exists(os.path.join(ROOT, "julia", f))
