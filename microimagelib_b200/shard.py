"""Time-point sharding across GPUs: independent units, no data-path collective.

The same round-robin rule as apps/spim_fusion_batch.cpp (MILB_SHARD=<rank>/<world>): the k-th time
point of the batch goes to rank k % world.  torch.distributed (gloo on CPU, NCCL on GPUs) is used
only for the barrier and for the max-over-ranks reduction of timings."""
from __future__ import annotations


def time_points(start: int, end: int, step: int):
    return list(range(start, end + 1, step))


def shard_time_points(start: int, end: int, step: int, rank: int, world: int):
    """Time points handled by `rank` out of `world` processes."""
    if not (0 <= rank < world) or step <= 0:
        raise ValueError("bad shard arguments")
    return [t for k, t in enumerate(time_points(start, end, step)) if k % world == rank]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (timing) over the process group; identity without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
