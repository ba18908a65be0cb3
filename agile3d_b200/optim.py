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
    """AdamW over the flat buffers.  One difference from torch.optim.AdamW: a parameter that received no gradient in a step
    (torch: ``p.grad is None`` -> skipped) has a zero gradient here and still gets its weight decay and moment decay;
    in the reference's training step every parameter of the model receives a gradient (engine.py:119-150)."""

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
    """Data-parallel gradient exchange of the training step: sum-then-average all-reduce (NCCL over NVLink / NVSwitch) of
    FlatAdamW's flat gradient buffer, bucket by bucket, OVERLAPPED with the backward.

    attach(model) makes the backward of the backbone write its gradients straight into the flat buffer, stage by stage in
    the order they are produced (decoder + head first, then block8/up7 ... block1/down1, the stem last), and hands every
    finished stage to this object: a stage is a contiguous slice of the flat buffer (parameters are registered in forward
    order), so its all-reduce is launched on a side stream - behind an event recorded on the compute stream - while the
    compute stream goes on with the data and weight gradients of the earlier stages.  finish() (called by all_reduce())
    makes the compute stream wait for the exchange before clip + AdamW.  Without attach() the whole buffer is exchanged
    after the backward in `n_buckets` slices (no overlap).  BatchNorm statistics stay local (the reference has no SyncBN)."""

    def __init__(self, opt: FlatAdamW, n_buckets: int = 6, group=None):
        self.opt, self.group = opt, group
        n = opt.flat_g.numel()
        edges = [round(i * n / n_buckets) for i in range(n_buckets + 1)]
        self.slices = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.stream = torch.cuda.Stream() if opt.flat_g.is_cuda else None
        self.model = None
        self._works, self._done_to = [], None
        self.overlap = True
        self.launched = []             # (start, end) of the slices exchanged during the last backward (diagnostics)

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    # ---- overlapped mode -------------------------------------------------------------------------------------
    def attach(self, model):
        """route the backbone's gradients through this object (model.grad_sink)"""
        by_param = {id(p): (off, k) for p, (off, k) in zip(self.opt.params, self.opt.offsets)}
        self.ranges = {}
        for name, p in model.named_parameters():
            if id(p) in by_param:
                self.ranges[name] = by_param[id(p)]
        offs = [self.ranges[n][0] for n in self.ranges if not n.startswith("backbone.")]
        self.tail_start = min(offs) if offs else self.opt.flat_g.numel()
        if any(self.ranges[n][0] >= self.tail_start for n in self.ranges if n.startswith("backbone.")):
            raise RuntimeError("GradBuckets.attach: backbone parameters must precede the decoder in parameter order")
        self.model = model
        model.grad_sink = self
        return self

    def detach(self):
        if self.model is not None:
            self.model.grad_sink = None
            self.model = None

    def write(self, grads):
        lo, hi = None, None
        for name, g in grads.items():
            off, k = self.ranges[name]
            self.opt.flat_g[off:off + k].copy_(g.reshape(-1))
            lo = off if lo is None else min(lo, off)
            hi = off + k if hi is None else max(hi, off + k)
        return lo, hi

    def _launch(self, a, b):
        self.launched.append((a, b))
        if self.world() == 1 or not self.overlap or b <= a:
            return
        if self.stream is None:                                # CPU process group (gloo tests): exchange in place, now
            dist.all_reduce(self.opt.flat_g[a:b], op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()                                            # the slice is complete on the compute stream here
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            self._works.append(dist.all_reduce(self.opt.flat_g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def tail_done(self):
        """everything from the first non-backbone parameter to the end of the buffer is final"""
        self.launched = []
        self._done_to = self.tail_start
        self._launch(self.tail_start, self.opt.flat_g.numel())

    def stage(self, grads):
        """gradients of one backbone stage (a contiguous slice ending where the previous one began)"""
        lo, _ = self.write(grads)
        if lo is not None and lo < self._done_to:
            self._launch(lo, self._done_to)
            self._done_to = lo

    def finish(self):
        w = self.world()
        if w == 1:
            self._works = []
            return
        g = self.opt.flat_g
        if not self.overlap or self.model is None or self._done_to is None:
            return self._all_reduce_after()
        if self._done_to > 0:                                  # whatever the stages did not cover (nothing, normally)
            self._launch(0, self._done_to)
        if self.stream is None:
            g.mul_(1.0 / w)
            self._works, self._done_to = [], None
            return
        with torch.cuda.stream(self.stream):
            for wk in self._works:
                wk.wait()
            g.mul_(1.0 / w)
        torch.cuda.current_stream().wait_stream(self.stream)
        self._works, self._done_to = [], None

    # ---- after-the-backward mode -------------------------------------------------------------------------------
    def _all_reduce_after(self):
        w = self.world()
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
        self._works, self._done_to = [], None

    def all_reduce(self):
        """Call after backward: returns when the compute stream may read the averaged gradients."""
        if self.world() == 1:
            self._works, self._done_to = [], None
            return
        self.finish()
