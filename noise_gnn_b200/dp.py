"""Data-parallel gradient exchange: one all-reduce of the flat fp32 gradient bucket per step.

The reference has no multi-device path; seed mini-batches shard naturally, so the only exchange
step is the gradient all-reduce: NCCL over NVLink on the GPUs, gloo in the CPU tests.  The bucket is a
single contiguous tensor (0.83 MB for the products model), so the collective is latency-bound and is issued once.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_mean_(bucket: torch.Tensor, group=None, world_size: int | None = None, prescaled: bool = False) -> float:
    """Sum `bucket` over the group in place.  Returns the factor the optimizer must still apply (1/world) unless
    `prescaled` (then the bucket is divided here)."""
    world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    if world <= 1:
        return 1.0
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if prescaled:
        bucket.div_(world)
        return 1.0
    return 1.0 / world
