"""Data-parallel leg of the hot path.

The reference wraps the model in single-process ``nn.DataParallel`` (wbwtab/main.py:323-327,
dorefa/main.py:302-306, iao/main.py:496-500): per step it re-broadcasts parameters, scatters the
batch, gathers logits and reduce-adds gradients onto GPU 0.  Here: one process per GPU
(torchrun), full replicas, rank-local BN statistics and observers (DataParallel does not
synchronise them either), and exactly ONE collective per step - an in-place NCCL all-reduce
(sum) over a flat fp32 gradient bucket, followed by a 1/world scale.  Parameter ``.grad``
tensors are views into the bucket, so there is no gather/scatter copy around the collective."""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        if not self.params:
            self.flat = None
            return
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def nbytes(self):
        return 0 if self.flat is None else self.flat.numel() * self.flat.element_size()

    def zero(self):
        """replacement for optimizer.zero_grad(): keeps the .grad views alive"""
        if self.flat is not None:
            self.flat.zero_()

    def all_reduce(self):
        """sum over ranks, then average - equals DataParallel's full-batch mean-loss gradient
        when every rank holds an equal shard."""
        if self.flat is None or not dist.is_available() or not dist.is_initialized():
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)


class FlatAdam:
    """Adam over flat parameter / gradient buckets: parameters and their ``.grad`` become views of two
    contiguous fp32 buffers, the update is one ``mnb_adam_step`` launch (same math as
    ``torch.optim.Adam(lr, betas, eps, weight_decay)`` with identical hyper-parameters in every
    group, which is what the reference's training scripts construct: wbwtab/main.py:331-339)."""

    def __init__(self, params, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        from . import _lib as L
        self._L = L
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        L.require_cuda(*self.params)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
                off += n
        self.bucket = FlatGradBucket(self.params, process_group)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)

    def zero_grad(self):
        self.bucket.zero()

    def all_reduce(self):
        self.bucket.all_reduce()

    def step(self):
        L = self._L
        self.step_count += 1
        L.check(L.load().mnb_adam_step(self.flat_p.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.flat_p.numel(), self.lr, self.betas[0],
                                       self.betas[1], self.eps, self.weight_decay, self.step_count, L.stream()),
                "adam_step")
