"""Data parallelism for the train step: shard drawings across ranks, one gradient all-reduce per step.

The reference trains with Lightning `strategy: ddp` (ref: configs/train_complete.yaml:18-21): one
process per GPU, each rank's loss is the mean over ITS non-PAD targets (ref models.py:221) and DDP
averages the gradients over ranks (mean of per-rank means -- not token weighted).  The model is 32 M
parameters, so there is nothing to shard but the batch and the only exchange is that all-reduce.

`GradAllReduce` keeps every parameter's .grad as a view into ONE flat fp32 buffer, so the exchange is a
single NCCL all-reduce over NVLink/NVSwitch (~130 MB) instead of DDP's 25 MB buckets; torch DDP also
works with the module (gradients reach the canonical nn.Parameters) and is what Lightning would use.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_samples: int, rank: int, world: int):
    """DistributedSampler-style strided shard: rank r gets samples r, r+world, ... (drop the ragged tail)."""
    per = n_samples // world
    return [rank + i * world for i in range(per)]


class GradAllReduce:
    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, device=p0.device, dtype=p0.dtype)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_grad(self):
        """Use instead of optimizer.zero_grad(set_to_none=True): the views must stay attached."""
        self.flat.zero_()

    def sync(self):
        """Average gradients over the ranks (call after backward, before optimizer.step)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)
