"""Data parallelism over the image batch: one process per GPU,
weights replicated, every image's episode independent.  Exactly two exchanges per
iteration, both through torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests):

  1. the advantage statistics {sum, sum of squares, count} (3 doubles) between the
     two phases of the fused loss, so ``standardize`` sees the global batch;
  2. ONE all-reduce of the flat gradient bucket the backward kernels wrote into
     (no packing copy), followed by the 1/world scale.
"""
from __future__ import annotations

from typing import Optional

import torch as th
import torch.distributed as dist


class DataParallelContext:
    def __init__(self, group: Optional["dist.ProcessGroup"] = None, enabled: Optional[bool] = None) -> None:
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.enabled = bool(enabled)
        self.group = group
        self.world_size = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    def shard(self, batch: th.Tensor) -> th.Tensor:
        """Rank r keeps images [r*B/G, (r+1)*B/G): all agents of an image stay together."""
        if not self.enabled:
            return batch
        n = batch.shape[0]
        if n % self.world_size != 0:
            raise RuntimeError(f"global batch {n} is not divisible by world size {self.world_size}")
        per = n // self.world_size
        return batch[self.rank * per: (self.rank + 1) * per]

    def all_reduce_stats(self, stats: th.Tensor) -> None:
        """stats[0:3] = {sum(adv), sum(adv^2), n}: summed over ranks in place."""
        if self.enabled:
            dist.all_reduce(stats[:3], op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_grads(self, flat_grads: th.Tensor) -> None:
        """One collective over the whole bucket, then average (the loss is a mean
        over each rank's equal shard, trainer.py:111)."""
        if self.enabled:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
            flat_grads.mul_(1.0 / self.world_size)

    def all_reduce_grads_sum(self, flat_grads: th.Tensor) -> None:
        """Sum only; the 1/world scale is folded into the fused Adam (grad_scale)."""
        if self.enabled:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)

    def broadcast_params(self, flat_params: th.Tensor, src: int = 0) -> None:
        if self.enabled:
            dist.broadcast(flat_params, src=src, group=self.group)
