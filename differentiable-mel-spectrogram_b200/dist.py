"""Data-parallel use of the layer: clips shard over ranks, lambda's gradient is the only exchange.

Frames never cross clips, so `x`, `y`, `dL/dy` and `dL/dx` stay rank-local; `lambd` is
replicated and its scalar gradient is summed with ONE all-reduce (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Under DistributedDataParallel nothing here is needed: `lambd` is an
ordinary parameter and DDP's bucket all-reduce covers it.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of the batch: the first n_clips % world ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_clips, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_lambda_grad(params: Iterable[torch.nn.Parameter], average: bool = False, group=None) -> None:
    """SUM (or mean) the .grad of the given lambd parameters over the process group, in place.

    Use SUM when each rank back-propagated the gradient of a global (summed) loss through its own
    shard — then the N-rank result equals the 1-rank result on the concatenated batch.
    All lambdas go out in one flat buffer so there is a single collective per step.
    """
    if not dist.is_available() or not dist.is_initialized():
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.stack([p.grad.detach().reshape(()) for p in ps]) if len(ps) > 1 else ps[0].grad.detach().reshape(1)
    flat = flat.contiguous()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    for i, p in enumerate(ps):
        p.grad.copy_(flat[i].reshape(p.grad.shape))


def broadcast_lambda(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Make lambd identical on all ranks (n_fft is derived from it, so it must be)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
