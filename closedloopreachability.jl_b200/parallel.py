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


def concat_shards(parts):
    """Results of consecutive shards back in global cell order."""
    return np.concatenate(list(parts), axis=0)
