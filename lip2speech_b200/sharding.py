"""Multi-GPU plumbing for the inference path: clips are independent, so the batch dimension is
partitioned contiguously across ranks with replicated weights and NO data-path collective.  torch.distributed is
used only for the barrier, the max-over-ranks of device timings and an optional gather of results."""
from __future__ import annotations

import torch
import torch.distributed as dist


def clip_range(total: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of `total` clips for `rank` of `world`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_clip_ids(mine: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather of per-rank clip ids (ragged shards padded to the largest); used by tests / result collection."""
    world = dist.get_world_size()
    cap = -(-total // world)
    buf = torch.full((cap,), -1, dtype=mine.dtype, device=mine.device)
    buf[: mine.numel()] = mine
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    cat = torch.cat(outs)
    return cat[cat >= 0]
