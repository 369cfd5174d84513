"""Episode sharding across ranks (one process per GPU).

Episodes are independent units: rank r owns a contiguous slice of the episode
batch, the head weights are replicated, and the only exchange is ONE all-reduce (sum) of the
flattened head-gradient bucket plus the [loss_sum, correct, episodes] scalars per optimizer step
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Feature gradients stay local.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of `total` episodes for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class HeadGradReducer:
    """Flattens the gradients of `params` into one fp32 bucket and sums it over ranks.

    The bucket is persistent; `reduce()` copies grads in, launches one all_reduce (optionally on a
    side stream so it overlaps whatever the caller enqueues next) and `finish()` writes the summed
    values back into the .grad tensors.

    With `grads_as_views=True` every `p.grad` IS a slice of the bucket (like DDP's gradient_as_bucket_view):
    backward accumulates straight into it, the all-reduce runs in place and nothing is copied in or out --
    2 x len(params) copy kernels per step less.  The caller then zeroes gradients with `zero()` (never with
    `optimizer.zero_grad(set_to_none=True)`, which would drop the views).
    """

    def __init__(self, params, group=None, side_stream: bool = True, grads_as_views: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.bucket = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = bool(grads_as_views)
        if self.views:
            off = 0
            for p in self.params:
                p.grad = self.bucket[off:off + p.numel()].view_as(p)
                off += p.numel()
        self.scalars = torch.zeros(3, dtype=torch.float64 if dev.type == "cpu" else torch.float32, device=dev)
        self.stream = torch.cuda.Stream(device=dev) if (side_stream and dev.type == "cuda") else None
        self._work = None

    @property
    def nbytes(self):
        return self.bucket.numel() * 4

    def zero(self):
        """Zero every gradient with one memset (bucket-view mode) or per tensor."""
        if self.views:
            self.bucket.zero_()
        else:
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()

    def reduce(self, loss_sum=None, correct=None, episodes=None):
        off = 0
        for p in ([] if self.views else self.params):
            n = p.numel()
            if p.grad is None:
                self.bucket[off:off + n].zero_()
            else:
                self.bucket[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if loss_sum is not None:
            # fill_/copy_ of device values only: safe inside CUDA-graph capture (no host->device scalar copies)
            for slot, val in ((0, loss_sum), (1, correct), (2, episodes)):
                dst = self.scalars[slot:slot + 1]
                if torch.is_tensor(val):
                    dst.copy_(val.reshape(1).to(dst.dtype))
                else:
                    dst.fill_(float(0 if val is None else val))
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(self.scalars, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.scalars, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        off = 0
        for p in ([] if self.views else self.params):
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.bucket[off:off + n].view_as(p))
            off += n
        return self.scalars


def sharded_step(compute, batch_size: int, params, reducer: HeadGradReducer | None = None, group=None):
    """Run `compute(lo, hi) -> (loss_sum, correct)` on this rank's episodes (it must leave the head
    gradients in .grad), then sum gradients and scalars over ranks.  Returns (loss_sum, correct,
    episodes) of the WHOLE batch, identical on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(batch_size, rank, world)
    loss_sum, correct = compute(lo, hi)
    reducer = reducer or HeadGradReducer(params, group, side_stream=False)
    reducer.reduce(loss_sum, correct, hi - lo)
    s = reducer.finish()
    return s[0], s[1], s[2]
