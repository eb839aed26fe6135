"""Multi-GPU plumbing: shots shard across ranks, counters are all-reduced.

One process per GPU (``torchrun``); shots are independent, so rank ``r`` of
``R`` takes the contiguous shot range ``[r*n//R, (r+1)*n//R)`` of the global
counter-based RNG stream and the only collective on the path is one
``all_reduce(SUM)`` of the ``int64`` counters per data point (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md section 8e.
"""
from __future__ import annotations


def world():
    """(rank, world_size) of the default process group, (0, 1) if none."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous share of ``n`` units for ``rank``: (offset, count)."""
    lo = rank * n // world_size
    hi = (rank + 1) * n // world_size
    return lo, hi - lo


def allreduce_counts(values, device=None):
    """Sum a list of non-negative ints over all ranks (identity for one rank)."""
    rank, size = world()
    if size == 1:
        return [int(v) for v in values]
    import torch
    import torch.distributed as dist
    if dist.get_backend() == 'nccl':        # NCCL reduces device tensors only
        # no device given: the one the engines pick by default (LOCAL_RANK), not torch's current device --
        # under torchrun that is cuda:0 on every rank unless the caller set it, and NCCL then sees one GPU twice
        import os
        dev = device if device is not None else f"cuda:{int(os.environ.get('LOCAL_RANK', torch.cuda.current_device())) % max(torch.cuda.device_count(), 1)}"
    else:
        dev = 'cpu'
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.cpu().tolist()]
