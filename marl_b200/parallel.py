"""Data-parallel protocol of the learner step.

Episodes never interact before the loss reduction, so the sampled batch shards along the episode axis
with NO data-path collective; the only exchange is ONE all-reduce (sum) per train step over the flat
buffer ``[d/dtheta of sum((mask*delta)^2)  |  sum((mask*delta)^2)  |  sum(mask)]``.  Every rank then
divides by the global sum(mask), clips by the same global norm and applies the same optimiser update to
its replica, so parameters stay bit-identical across ranks without a broadcast.
"""
from __future__ import annotations


def shard_bounds(n_episodes, world_size, rank):
    """Contiguous episode shard of rank `rank`: [lo, hi).  The batch size must divide evenly."""
    per = n_episodes // world_size
    if per * world_size != n_episodes:
        raise ValueError("data-parallel training needs the batch size to be a multiple of the world size")
    return rank * per, (rank + 1) * per


def allreduce_flat(grad_full, dist, group=None):
    """Sum the flat [grad | loss_sum | mask_sum] buffer over the ranks, in place (NCCL on GPUs, gloo in tests)."""
    dist.all_reduce(grad_full, op=dist.ReduceOp.SUM, group=group)
    return grad_full
