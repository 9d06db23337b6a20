"""Data parallelism for the train step: shard drawings across ranks, one gradient all-reduce per step.

The reference trains with Lightning `strategy: ddp` (ref: configs/train_complete.yaml:18-21): one
process per GPU, each rank's loss is the mean over ITS non-PAD targets (ref models.py:221) and DDP
averages the gradients over ranks (mean of per-rank means -- not token weighted).  The model is 32 M
parameters, so there is nothing to shard but the batch and the only exchange is that all-reduce.

`GradAllReduce` exchanges ALL gradients of a step with ONE NCCL all-reduce (average) over NVLink/NVSwitch (~130 MB):
after backward the per-parameter gradients (which autograd hands out as fresh tensors -- no pre-existing `.grad`, so no
AccumulateGrad add kernels) are gathered into one flat buffer by a single concatenation, reduced in place with
`ReduceOp.AVG` (no separate division), and `.grad` of every parameter becomes a view of that buffer, which the optimizer
reads in place.  Round 1 kept `.grad` as pre-existing views of a flat buffer instead: autograd then issued one extra add
kernel per parameter (~270 launches, +1.1 ms per step at N=2 -- most of the round-1 scaling loss) plus a 130 MB `div_`.
torch DDP also works with the module (gradients reach the canonical nn.Parameters; tests/test_trainer_boundary.py) and is
what Lightning would use.
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
        self.avg = dist.is_initialized() and dist.get_backend(group) == 'nccl'       # gloo has no AVG
        self.flat = None

    def zero_grad(self):
        """Same as optimizer.zero_grad(set_to_none=True): autograd then STEALS each gradient instead of adding into it."""
        for p in self.params:
            p.grad = None

    def sync(self):
        """Average gradients over the ranks (call after backward, before optimizer.step)."""
        ps = [p for p in self.params if p.grad is not None]
        if self.world <= 1 or not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])          # one gather kernel, 130 MB
        if self.avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad = flat[off:off + n].view_as(p)
            off += n
        self.flat = flat
