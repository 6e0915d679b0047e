"""Host-side data-parallel plumbing (one process per GPU, torch.distributed): batch sharding, the single gradient
all-reduce of a training step, max-over-ranks timing.  Pure torch.distributed calls - the same code runs over NCCL
on the B200 box and over gloo in the CPU tests."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [start, start+count) of `total` units for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """The ONE collective of a training step: in-place sum of the flat gradient buffer over all ranks.  Callers
    pre-scale the loss gradient by 1/world (up_mse_fwd_bwd's gscale), so the sum is the global-batch mean gradient."""
    if world()[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(values: Sequence[float], device) -> Sequence[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
