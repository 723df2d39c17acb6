"""Ray-sharded multi-GPU mapping (not in the reference, which is single-GPU):
one process per GPU, parameters replicated, every rank renders its own slice of the keyframe-ray
batch, and two exchanges per iteration make all ranks differentiate the SAME global-batch loss:

  1. all-reduce of the 8 int32 loss normalisers (ray / band counts) -> `norm` for eslam_loss_backward
  2. all-reduce (sum, fp32) of the gradient arena (12 planes + decoders, one contiguous buffer) and of
     the [frames,12] pose-gradient block, over NCCL (NVLink 5 / NVSwitch), then the identical Adam step.

The same code runs on CPU tensors with the gloo backend for the world_size-2 tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class MappingExchange:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._norm = None

    def reduce_counters(self, counters: torch.Tensor) -> torch.Tensor:
        if self._norm is None or self._norm.device != counters.device:
            self._norm = torch.empty_like(counters)
        self._norm.copy_(counters)
        if self.world > 1:
            dist.all_reduce(self._norm, op=dist.ReduceOp.SUM, group=self.group)
        return self._norm

    def reduce_grads(self, grad_arena: torch.Tensor, pose_grad=None, loss_acc=None) -> None:
        if self.world == 1:
            return
        dist.all_reduce(grad_arena, op=dist.ReduceOp.SUM, group=self.group)
        if pose_grad is not None:
            dist.all_reduce(pose_grad, op=dist.ReduceOp.SUM, group=self.group)
        if loss_acc is not None:
            dist.all_reduce(loss_acc, op=dist.ReduceOp.SUM, group=self.group)


def shard_range(total: int, rank: int, world: int):
    """Contiguous [start, count) slice of `total` independent units for `rank` (mesh voxel blocks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)
