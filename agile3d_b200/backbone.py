"""Res16UNet34C sparse-voxel backbone on the B200 library.

Mirrors the reference graph (models/res16unet.py:26-295, models/resnet.py:96-149,
models/modules/resnet_block.py:7-64, models/modules/common.py:125-188) and its state_dict layout
 so reference checkpoints load with ``load_state_dict``; the arithmetic is in
csrc/ (hash tables, kernel maps, fused sparse convolutions).

Eval-mode execution fuses every MinkowskiBatchNorm into the producing convolution's epilogue (scale/shift),
ReLU and the residual add likewise, and `me.cat` is free: producers write straight into a channel slice of the
concatenated buffer.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops

PLANES = (32, 64, 128, 256, 256, 128, 96, 96)     # Res16UNet34C  (models/res16unet.py:371-372)
LAYERS = (2, 3, 4, 6, 2, 2, 2, 2)                 # Res16UNet34   (models/res16unet.py:310)
INIT_DIM = 32
BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------ coordinate maps
class CoordinateMaps:
    """What MinkowskiEngine keeps in its CoordinateManager for one input tensor: 5 coordinate levels with their
    hash tables, and the kernel maps of every (level, kernel) pair the U-Net uses.  Built once per
    SparseTensor on the GPU and cached on it."""

    def __init__(self, coords: torch.Tensor, count_pairs: bool = False):
        if coords.dtype != torch.int32 or coords.dim() != 2 or coords.shape[1] != 4 or not coords.is_contiguous():
            raise ValueError("coords must be a contiguous int32 [N,4] tensor")
        self.coords = [coords]
        table, cap, status = ops.hash_build(coords)
        self.tables, self.caps, self.parents = [table], [cap], []
        self._status = status
        for lvl in range(4):                                        # tensor strides 2, 4, 8, 16
            c, t, cp, par = ops.downsample(self.coords[lvl], 2 << lvl)
            self.coords.append(c)
            self.tables.append(t)
            self.caps.append(cp)
            self.parents.append(par)
        # the downsample calls synchronised, so the status words are ready
        dup, oor = (int(v) for v in status.tolist())
        if dup or oor:
            raise ValueError(f"SparseTensor coordinates: {dup} duplicate rows, {oor} rows outside +-32767 "
                             "(run sparse_quantize first)")
        self.sizes = [int(c.shape[0]) for c in self.coords]
        self.pair_counts = {}
        self.k3, self.down, self.up = [], [], []
        for lvl in range(5):                                        # 3x3x3 maps, shared by all blocks of a level
            r = ops.kernel_map(self.coords[lvl], self.tables[lvl], self.caps[lvl], 3, 1 << lvl, 1, count_pairs)
            self.k3.append(self._keep(("k3", lvl), r, count_pairs))
        for lvl in range(4):                                        # kernel-2 stride-2 maps and their transposes
            r = ops.kernel_map(self.coords[lvl + 1], self.tables[lvl], self.caps[lvl], 2, 1 << lvl, 1, count_pairs)
            self.down.append(self._keep(("down", lvl), r, count_pairs))
            self.up.append(ops.kernel_map_transposed(self.coords[lvl], self.parents[lvl], 1 << lvl))
            if count_pairs:
                self.pair_counts[("up", lvl)] = self.sizes[lvl]

    def _keep(self, key, r, count_pairs):
        if count_pairs:
            self.pair_counts[key] = int(r[1].sum().item())
            return r[0]
        return r


# ------------------------------------------------------------------------------------------ parameter holders
class SparseConv(nn.Module):
    """Parameter layout of MinkowskiConvolution(Transpose): ``kernel`` [K,cin,cout] ([cin,cout] for 1x1), ``bias`` [1,cout]."""

    def __init__(self, cin, cout, ksize, stride=1, bias=False, transposed=False):
        super().__init__()
        self.cin, self.cout, self.ksize, self.stride, self.transposed = cin, cout, ksize, stride, transposed
        K = ksize ** 3
        shape = (cin, cout) if (K == 1 and stride == 1) else (K, cin, cout)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, cout)) if bias else None
        s = 1.0 / math.sqrt((cout if transposed else cin) * K)
        with torch.no_grad():
            self.kernel.uniform_(-s, s)
            if self.bias is not None:
                self.bias.uniform_(-s, s)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        k = state_dict.get(prefix + "kernel")
        if k is not None and k.dim() == 3 and self.kernel.dim() == 2 and k.shape[0] == 1:
            state_dict[prefix + "kernel"] = k[0]                    # accept [1,cin,cout] for 1x1 kernels
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class SparseBatchNorm(nn.Module):
    """MinkowskiBatchNorm: parameters live under ``.bn.*`` (models/modules/common.py:20-22)."""

    def __init__(self, c, momentum=0.1):
        super().__init__()
        self.bn = nn.BatchNorm1d(c, eps=BN_EPS, momentum=momentum)

    def folded(self):
        bn = self.bn
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        return scale.float().contiguous(), shift.float().contiguous()


class BasicBlock(nn.Module):
    def __init__(self, cin, planes, downsample=None):
        super().__init__()
        self.conv1 = SparseConv(cin, planes, 3)
        self.norm1 = SparseBatchNorm(planes, 0.1)      # blocks keep the default momentum
        self.conv2 = SparseConv(planes, planes, 3)
        self.norm2 = SparseBatchNorm(planes, 0.1)
        self.downsample = downsample


def _stage(cin, planes, n, momentum):
    ds = None
    if cin != planes:
        ds = nn.Sequential(SparseConv(cin, planes, 1), SparseBatchNorm(planes, momentum))
    return nn.Sequential(BasicBlock(cin, planes, ds), *[BasicBlock(planes, planes) for _ in range(n - 1)])


_ENC = ("1p1", "2p2", "3p4", "4p8")
_DEC = ("4p16", "5p8", "6p4", "7p2")


class Res16UNet34C(nn.Module):
    """build_backbone(args) of the reference (models/backbone.py:5-7): Res16UNet34C(3, 20, args, out_fpn=True)."""

    PLANES = PLANES

    def __init__(self, in_channels=3, bn_momentum=0.02, conv1_kernel_size=5):
        super().__init__()
        if in_channels != 3:
            raise ValueError("the stem kernel is specialised for 3 input channels (RGB), as build_backbone uses")
        m, P, L = bn_momentum, PLANES, LAYERS
        self.conv1_kernel_size = conv1_kernel_size
        self.conv0p1s1 = SparseConv(in_channels, INIT_DIM, conv1_kernel_size)
        self.bn0 = SparseBatchNorm(INIT_DIM, m)
        c = INIT_DIM
        for i, tag in enumerate(_ENC):
            setattr(self, f"conv{tag}s2", SparseConv(c, c, 2, stride=2))
            setattr(self, f"bn{i + 1}", SparseBatchNorm(c, m))
            setattr(self, f"block{i + 1}", _stage(c, P[i], L[i], m))
            c = P[i]
        skips = (P[2], P[1], P[0], INIT_DIM)
        for j, tag in enumerate(_DEC):
            setattr(self, f"convtr{tag}s2", SparseConv(c, P[4 + j], 2, stride=2, transposed=True))
            setattr(self, f"bntr{4 + j}", SparseBatchNorm(P[4 + j], m))
            setattr(self, f"block{5 + j}", _stage(P[4 + j] + skips[j], P[4 + j], L[4 + j], m))
            c = P[4 + j]
        self.algo = ops.ALGO_AUTO
        self.split_rows = True      # keep activations as bf16 hi/lo pair rows between tensor-core layers
        self._fold_cache = None

    # -- per-checkpoint constants (folded BatchNorm scale/shift, bf16 hi/lo weight images for the tensor-core
    #    path), recomputed only when a parameter/buffer changed
    def _folded(self):
        key = (self.algo,) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._fold_cache is None or self._fold_cache[0] != key:
            table = {name: mod.folded() for name, mod in self.named_modules() if isinstance(mod, SparseBatchNorm)}
            if self.algo != ops.ALGO_SIMT:
                for name, mod in self.named_modules():
                    if isinstance(mod, SparseConv) and mod.cin % 32 == 0:
                        table["tc:" + name] = ops.prepare_tc_weight(mod.kernel)
            self._fold_cache = (key, table)
        return self._fold_cache[1]

    def _conv(self, name, conv, x, nbr, out, scale, shift, fold, residual=None, relu=False):
        # tensor-core mode keeps every backbone activation as bf16 hi/lo pair rows ("split", same bytes as fp32):
        # producers split once in their epilogue, consumers gather with cp.async and no conversion work
        sp = self.split_rows and self.algo != ops.ALGO_SIMT
        return ops.spconv_fwd(x, nbr, conv.kernel, out, scale, shift, residual=residual, relu=relu, algo=self.algo,
                              weight_tc=fold.get("tc:" + name), in_split=sp, out_split=sp, res_split=sp)

    def _block(self, prefix, blk, x, nbr, fold, out=None):
        n = x.shape[0]
        dev = x.device
        planes = blk.conv1.cout
        s1, b1 = fold[prefix + ".norm1"]
        t = torch.empty((n, planes), dtype=torch.float32, device=dev)
        self._conv(prefix + ".conv1", blk.conv1, x, nbr, t, s1, b1, fold, relu=True)
        if blk.downsample is not None:
            sd, bd = fold[prefix + ".downsample.1"]
            res = torch.empty((n, planes), dtype=torch.float32, device=dev)
            self._conv(prefix + ".downsample.0", blk.downsample[0], x, None, res, sd, bd, fold)
        else:
            res = x
        s2, b2 = fold[prefix + ".norm2"]
        if out is None:
            out = torch.empty((n, planes), dtype=torch.float32, device=dev)
        self._conv(prefix + ".conv2", blk.conv2, t, nbr, out, s2, b2, fold, residual=res, relu=True)
        return out

    def _stage_fwd(self, name, x, nbr, fold, final_out=None):
        blocks = getattr(self, name)
        for i, blk in enumerate(blocks):
            x = self._block(f"{name}.{i}", blk, x, nbr, fold, out=final_out if i == len(blocks) - 1 else None)
        return x

    @torch.no_grad()
    def forward(self, st):
        """st: agile3d_b200.SparseTensor -> (features [N0, 96], [5 feature maps], CoordinateMaps)."""
        if self.training:
            raise NotImplementedError("train-mode (batch-statistics BatchNorm + backward) is not built yet; "
                                      "call model.eval() — see DESIGN.md 'out of scope this round'")
        if st.maps is None:
            st.maps = CoordinateMaps(st.C)
        maps, fold, dev = st.maps, self._folded(), st.F.device
        N, P = maps.sizes, PLANES
        f32 = dict(dtype=torch.float32, device=dev)
        # concat buffers: [upsampled | skip]  (me.cat(out, skip), models/res16unet.py:257,267,277,287)
        skip_c = (INIT_DIM, P[0], P[1], P[2])
        up_c = (P[7], P[6], P[5], P[4])
        cat = [torch.empty((N[l], up_c[l] + skip_c[l]), **f32) for l in range(4)]
        # stem: conv0p1s1 + bn0 + relu -> skip slice of the level-0 concat buffer
        s0, b0 = fold["bn0"]
        ops.stem_conv_fwd(maps.coords[0], st.F, maps.tables[0], maps.caps[0], self.conv1_kernel_size,
                          self.conv0p1s1.kernel, cat[0][:, up_c[0]:], s0, b0, relu=True,
                          out_split=self.split_rows and self.algo != ops.ALGO_SIMT)
        y = cat[0][:, up_c[0]:]
        for i, tag in enumerate(_ENC):                                  # encoder
            conv = getattr(self, f"conv{tag}s2")
            s, b = fold[f"bn{i + 1}"]
            d = torch.empty((N[i + 1], conv.cout), **f32)
            self._conv(f"conv{tag}s2", conv, y, maps.down[i], d, s, b, fold, relu=True)
            dst = cat[i + 1][:, up_c[i + 1]:] if i < 3 else None       # block output doubles as the skip
            y = self._stage_fwd(f"block{i + 1}", d, maps.k3[i + 1], fold, final_out=dst)
        fmaps = [y]
        for j, tag in enumerate(_DEC):                                  # decoder
            lvl = 3 - j
            conv = getattr(self, f"convtr{tag}s2")
            s, b = fold[f"bntr{4 + j}"]
            self._conv(f"convtr{tag}s2", conv, y, maps.up[lvl], cat[lvl][:, :up_c[lvl]], s, b, fold, relu=True)
            y = self._stage_fwd(f"block{5 + j}", cat[lvl], maps.k3[lvl], fold)
            fmaps.append(y)
        return y, fmaps, maps
