"""Deterministic synthetic checkpoints in the reference's state_dict layout.

There is no network for the pretrained checkpoint, so tests and bench.py use random-init weights of
the reference architecture.  `synth_state_dict` fills any {name: shape} mapping (taken from
``model.state_dict()`` of either the reference model or ours) with values that depend only on
(seed, name), so the reference model (golden generation) and this package get identical weights.
"""
from __future__ import annotations

import math
import zlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _k_eff(K: int) -> float:
    # expected number of occupied kernel offsets on surface-like scans
    return {1: 1.0, 8: 4.0, 27: 12.0, 125: 33.0}.get(K, float(K))


def synth_state_dict(shapes, seed: int = 0, dtype=torch.float32):
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.tensor(100, dtype=torch.long)
            continue
        if leaf == "kernel":
            K = shape[0] if len(shape) == 3 else 1
            cin = shape[-2]
            tr = "convtr" in name
            gain = 0.5 if (".conv2." in name or "downsample" in name) else 1.0   # keep residual sums O(1)
            if name.startswith("lin_squeeze_head"):
                gain = 40.0                                                    # decoder inputs of unit scale
            std = math.sqrt(gain / (cin * (1.0 if tr else _k_eff(K))))
            t = torch.randn(shape, generator=g) * std
        elif ".bn." in name or name.endswith(("running_mean", "running_var")):
            if leaf == "running_var":
                t = 0.8 + 0.4 * torch.rand(shape, generator=g)
            elif leaf == "weight":
                t = 0.8 + 0.4 * torch.rand(shape, generator=g)
            else:                                   # bias, running_mean
                t = 0.05 * torch.randn(shape, generator=g)
        elif "norm" in name and leaf == "weight":   # LayerNorm gains
            t = 0.9 + 0.2 * torch.rand(shape, generator=g)
        elif "norm" in name and leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g)
        elif name in ("bg_query_feat.weight", "bg_query_pos.weight", "pos_enc.gauss_B"):
            t = torch.randn(shape, generator=g)
        elif len(shape) >= 2:                       # xavier-uniform for every other matrix
            fan_out, fan_in = shape[0], shape[1]
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        else:                                       # biases
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def default_args(**overrides):
    """argparse-style namespace with the model-relevant defaults of the reference (main.py:34-58)."""
    from types import SimpleNamespace

    ns = SimpleNamespace(
        dialations=[1, 1, 1, 1], conv1_kernel_size=5, bn_momentum=0.02, voxel_size=0.05,
        hidden_dim=128, dim_feedforward=1024, num_heads=8, num_decoders=3, num_bg_queries=10,
        dropout=0.0, pre_norm=False, normalize_pos_enc=True, positional_encoding_type="fourier",
        gauss_scale=1.0, hlevels=[4], shared_decoder=False,
        losses=["bce", "dice"], bce_loss_coef=1.0, dice_loss_coef=2.0, aux=True,
    )
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns
