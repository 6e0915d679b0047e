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


def make_buckets(items: Sequence[Tuple[int, int, int]], bucket_bytes: int, n_ops: int, elem_bytes: int = 4):
    """Cut a flat gradient buffer into contiguous all-reduce buckets.

    `items`: (offset, numel, ready) per parameter in LAYOUT order - `ready` = index into the backward op list after
    which that gradient is final.  Returns [(lo, hi, ready)]: elements [lo, hi) may be reduced once `ready` backward
    ops have run; `ready` is non-decreasing over the buckets and the last one equals n_ops (whole backward done)."""
    buckets, cur_start, cur_ready, cur_bytes, end = [], 0, 0, 0, 0
    for off, numel, ready in items:
        cur_ready = max(cur_ready, ready)
        cur_bytes += numel * elem_bytes
        end = off + numel
        if cur_bytes >= bucket_bytes:
            buckets.append([cur_start, end, cur_ready])
            cur_start, cur_bytes = end, 0
    if cur_bytes > 0 or not buckets:
        buckets.append([cur_start, end, cur_ready])
    run = 0
    for b in buckets:       # the layout only approximates the completion order: make the segment ends monotonic
        run = max(run, b[2])
        b[2] = run
    buckets[-1][2] = n_ops
    return [tuple(b) for b in buckets]


def allreduce_buckets_(flat: torch.Tensor, buckets) -> torch.Tensor:
    """Bucket-by-bucket in-place sum (what TrainStep overlaps with the backward); == allreduce_sum_(flat)."""
    if world()[1] > 1:
        for lo, hi, _ready in buckets:
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(values: Sequence[float], device) -> Sequence[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
