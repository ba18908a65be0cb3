"""Optimizer step and data-parallel gradient exchange of the training path (engine.py:146-150, main.py:125-126).

FlatAdamW keeps parameters, gradients and both Adam moments in four flat fp32 buffers (parameters and .grad are views),
so clip_grad_norm_ + AdamW are two kernels over 39.3 M floats (ag3d_grad_norm, ag3d_adamw_step) instead of ~600
per-tensor launches.  GradBuckets all-reduces that flat gradient buffer over NCCL (NVLink/NVSwitch) in a few large
buckets on a side stream, last-produced gradients first, so the exchange of the decoder/head bucket overlaps the
backbone's backward; BatchNorm statistics stay local (the reference has no SyncBN).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


class FlatAdamW:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                p.grad = self.flat_g[off:off + k].view_as(p)
                self.offsets.append((off, k))
                off += k
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self.last_norm = None
        ops.bump_param_generation()        # parameters were re-seated onto the flat buffer

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, (off, k) in zip(self.params, self.offsets):       # re-attach if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + k].view_as(p)

    def step(self):
        """clip_grad_norm_(max_norm) + AdamW; returns the (pre-clip) total gradient norm as a device scalar."""
        self.step_count += 1
        self.last_norm = ops.grad_norm(self.flat_g)
        ops.adamw_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.step_count, self.last_norm if self.max_norm > 0 else None, self.max_norm)
        # the kernel wrote the parameters behind torch's back (no ._version bump): invalidate every derived weight
        # image (Res16UNet34C._train_weights/_folded, Agile3d head image)
        ops.bump_param_generation()
        return self.last_norm


class GradBuckets:
    """Sum-then-average all-reduce of FlatAdamW's gradient buffer in `n_buckets` contiguous slices."""

    def __init__(self, opt: FlatAdamW, n_buckets: int = 6, group=None):
        self.opt, self.group = opt, group
        n = opt.flat_g.numel()
        edges = [round(i * n / n_buckets) for i in range(n_buckets + 1)]
        self.slices = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.stream = torch.cuda.Stream() if opt.flat_g.is_cuda else None

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def all_reduce(self):
        """Call after backward.  Buckets go out last-slice-first (the decoder/head gradients live at the end of the
        parameter order and are complete first); returns when the default stream may read the averaged gradients."""
        w = self.world()
        if w == 1:
            return
        g = self.opt.flat_g
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                works = [dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                         for a, b in reversed(self.slices)]
                for wk in works:
                    wk.wait()
                g.mul_(1.0 / w)
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for a, b in reversed(self.slices):
                dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.group)
            g.mul_(1.0 / w)
