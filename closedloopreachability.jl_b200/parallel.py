"""Multi-GPU plumbing: one process per GPU, cells / trajectories sharded by contiguous index range.

The hot path has no exchange step: every rank runs the same kernels on its
own range with replicated weights and NO collective in the loop.  The only communication is the
final gather of the per-GPU output-box hulls (2 m doubles per rank) -- an all_gather over NCCL
(NVLink) on device tensors, or gloo on CPU tensors in the tests -- followed by a min / max.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """[begin, begin+count) of rank's contiguous share of n units (product-order cell ids)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    b = (n * rank) // world
    e = (n * (rank + 1)) // world
    return b, e - b


def gather_hull(hull_lo, hull_hi, group=None):
    """Union hull over all ranks.  Tensors stay on their device (NCCL) or on the CPU (gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return hull_lo, hull_hi
    w = dist.get_world_size(group)
    mine = torch.stack([torch.as_tensor(hull_lo), torch.as_tensor(hull_hi)])
    buf = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(buf, mine, group=group)
    allh = torch.stack(buf)
    return allh[:, 0].amin(dim=0), allh[:, 1].amax(dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """MAX reduction of a per-rank scalar (timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device=None, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def gather_scalars(value: float, device=None, group=None):
    """The per-rank values of a scalar, in rank order, on every rank (per-rank timings in the bench line)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(value)]
    w = dist.get_world_size(group)
    mine = torch.tensor([float(value)], dtype=torch.float64, device=device)
    buf = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(buf, mine, group=group)
    return [float(b.item()) for b in buf]


def bounds_from_costs(edges, cost_per_cell, world: int):
    """Contiguous shard boundaries of equal estimated COST: `edges` (B + 1 ascending cell ids) delimit B blocks whose
    cells cost `cost_per_cell[i]` each; returns world + 1 cell ids (first = edges[0], last = edges[-1]) such that every
    shard carries 1 / world of the total cost (piecewise-constant cost density, boundaries interpolated inside a block)."""
    edges = [int(e) for e in edges]
    B = len(edges) - 1
    if B < 1 or len(cost_per_cell) != B or world < 1:
        raise ValueError("bad cost table")
    w = [max(float(c), 0.0) * (edges[i + 1] - edges[i]) for i, c in enumerate(cost_per_cell)]
    total = sum(w)
    if total <= 0.0:
        return [edges[0] + ((edges[-1] - edges[0]) * r) // world for r in range(world + 1)]
    bounds = [edges[0]]
    acc, i = 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while i < B - 1 and acc + w[i] < target:
            acc += w[i]
            i += 1
        frac = (target - acc) / w[i] if w[i] > 0 else 0.0
        b = edges[i] + int(round(min(max(frac, 0.0), 1.0) * (edges[i + 1] - edges[i])))
        bounds.append(max(bounds[-1], min(b, edges[-1])))
    bounds.append(edges[-1])
    return bounds


def balanced_bounds(ctx, dnet, grid, world: int, blocks_per_rank: int = 16, minis: int = 8, mini_cells: int = 64,
                    pre=None, post=None, base_cost: float = 0.0):
    """Shard boundaries weighted by a pilot's per-cell cost: the id range is cut into
    blocks_per_rank * world blocks, `minis` short contiguous ranges of every block go through the statistics
    instantiation of the DeepZ kernel, and a cell's cost is modelled as base_cost + its algorithmic FLOPs.  Every rank
    derives the same boundaries (the sample and the statistics are deterministic).  Returns (bounds, cost per block)."""
    n = len(grid)
    B = max(1, blocks_per_rank * world)
    edges = [(n * i) // B for i in range(B + 1)]
    cost = []
    for i in range(B):
        b0, b1 = edges[i], edges[i + 1]
        f, c = 0.0, 0
        for k in range(minis):
            s = b0 + ((2 * k + 1) * (b1 - b0)) // (2 * minis) + (k * 7919) % 997
            s = min(max(b0, s), max(b0, b1 - mini_cells))
            cnt = min(mini_cells, b1 - s)
            if cnt <= 0:
                continue
            r = ctx.deepz_boxgrid(dnet, grid.partition, grid.centers, grid.radii, cell_begin=s, cell_count=cnt, pre=pre,
                                  post=post, stats=True)
            f += r["stats"]["flops"]
            c += cnt
        cost.append(base_cost + (f / c if c else 0.0))
    return bounds_from_costs(edges, cost, world), cost


def concat_shards(parts):
    """Results of consecutive shards back in global cell order."""
    return np.concatenate(list(parts), axis=0)
