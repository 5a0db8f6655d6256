"""Sharding of a batch across the GPUs of one box: by matrix index, no collective on the data path
(SURVEY.md section 8e). One process per GPU (torchrun) or one host thread over per-device queues
(magma_b200_dgetrf_batched_mgpu / magma_b200_dgesv_batched_mgpu in the C ABI).

Host-side logic only -- importing this module needs neither a GPU nor the oracle.
"""
from __future__ import annotations

import numpy as np


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank`; shard sizes differ by at most one matrix."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lu_cost(m, n):
    """FLOPS_DGETRF(m, n) of testing/flops.h:79-84 -- the load measure for variable-size batches."""
    m = np.asarray(m, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    k = np.minimum(m, n)
    big = np.maximum(m, n)
    mul = 0.5 * k * (k * (big - k / 3.0 - 1.0) + big) + 2.0 * k / 3.0
    add = 0.5 * k * (k * (big - k / 3.0) - big) + k / 6.0
    return mul + add


def lpt_partition(m, n, world: int) -> list[np.ndarray]:
    """Variable-size batches: longest-processing-time-first assignment of matrices to ranks by
    LU cost. Returns, per rank, the (sorted) indices it owns. Deterministic."""
    cost = lu_cost(m, n)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(cost), dtype=np.int64)
    for i in order:
        g = int(np.argmin(load))
        owner[i] = g
        load[g] += cost[i]
    return [np.sort(np.nonzero(owner == g)[0]) for g in range(world)]


def imbalance(m, n, parts) -> float:
    """max rank load / mean rank load for a partition (1.0 = perfect)."""
    cost = lu_cost(m, n)
    loads = np.array([cost[p].sum() for p in parts])
    return float(loads.max() / loads.mean()) if loads.mean() > 0 else 1.0


def max_over_ranks(x: float, device=None) -> float:
    """Timed regions are reported as the maximum over ranks (the job is done when the last GPU is)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
