"""Batch sharding of the reconstruction loss across the GPUs of one node (one process per GPU).

Every image is independent (per-image K, pose, depth, sources); the only cross-sample coupling of
loss_functions.py:13 / loss_functions_sfm.py:33 is the `mean`, whose denominator B*C*H*W is known in
advance.  So each rank runs the same fused kernel on its contiguous slice of the batch and

    loss_global = sum_r (B_r / B) * loss_r ,     d loss_global / d x_r = (B_r / B) * d loss_r / d x_r

No data-path collective is needed; the loss terms are all-reduced (<= 16 floats) for logging only.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous [begin, end) slice of the batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def local_weight(global_batch: int, rank: int, world: int) -> float:
    """Factor that turns a rank-local mean loss (and its gradients) into its share of the global mean."""
    b, e = shard_range(global_batch, rank, world)
    return (e - b) / float(global_batch)


def all_reduce_terms(local_terms: torch.Tensor, weight: float, group=None, async_op: bool = False):
    """Global loss terms = sum over ranks of weight_r * local_terms_r (one tiny all-reduce, NCCL or gloo)."""
    buf = local_terms.detach() * weight
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return (buf, work) if async_op else buf
    return (buf, None) if async_op else buf
