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
import os

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

    # levels whose rows are re-ordered by neighbour pattern.  Measured at 8 x 150k voxels (profiles/r02_layer_times_*): level 0
    # gains 1.7 ms of convolution time for 0.5 ms of sorting / map rewriting; level 1 gains 0.34 ms for 0.25 ms, level 2 nothing
    REORDER_LEVELS = (0,)

    def __init__(self, coords: torch.Tensor, count_pairs: bool = False, want_offsets: bool = False, reorder: bool = False):
        if coords.dtype != torch.int32 or coords.dim() != 2 or coords.shape[1] != 4 or not coords.is_contiguous():
            raise ValueError("coords must be a contiguous int32 [N,4] tensor")
        # 5 coordinate levels (tensor strides 1 .. 16), their hash tables and the scene row ranges: the level sizes stay
        # on the device between the levels, the host reads everything back once (ops.build_levels)
        self.coords, self.tables, self.caps, self.parents, (dup, oor), self.offsets = \
            ops.build_levels(coords, 4, want_offsets)
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
        # bricks: full-resolution rows of the 4x4x4 cells under every tensor-stride-4 voxel (level 2); the stem finds its
        # 5x5x5 neighbours through them (8 probes of the level-2 table instead of 125 of the level-0 table)
        self.bricks = (self.tables[2], self.caps[2], ops.brick_rows(self.coords[0], self.parents[0], self.parents[1],
                                                                    self.sizes[2]))
        self.perm, self.inv = [None] * 5, [None] * 5
        if reorder:
            self._reorder()

    def _reorder(self):
        """Internal row order (csrc/coords.cu: ag3d_row_order): rows of a level sorted, scene by scene, by the pattern of
        their existing 3x3x3 neighbours, so that the rows of a 128-row tile agree on the absent offsets and the dense-tile
        convolution skips those stages.  perm[l][new] = old row, inv[l][old] = new row; every neighbour table is rewritten
        into the new numbering, coords[l] follows.  The hash tables keep the OLD row ids: the level-0 table is how the stem
        finds the input features, which stay in the caller's order."""
        for l in self.REORDER_LEVELS:
            if self.sizes[l] >= 256:
                self.perm[l], self.inv[l] = ops.row_order(self.k3[l], self.coords[l])
        for l in range(5):
            if self.perm[l] is not None:
                self.k3[l] = ops.permute_map(self.k3[l], self.perm[l], self.inv[l])
                self.coords[l] = ops.gather_rows(self.coords[l], self.perm[l])
        for l in range(4):
            if self.perm[l] is not None or self.perm[l + 1] is not None:
                self.down[l] = ops.permute_map(self.down[l], self.perm[l + 1], self.inv[l])      # out: level l+1, in: level l
                self.up[l] = ops.permute_map(self.up[l], self.perm[l], self.inv[l + 1])          # out: level l, in: level l+1

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
        self.reorder_rows = True    # internal row order by neighbour pattern (CoordinateMaps._reorder); callers never see it
        # train: weight gradients on a side stream under the data-gradient chain.  Opt-in (AG3D_WGRAD_STREAM=1): measured
        # 82.13 vs 82.20 ms/step at 4 x 150k voxels - the training step is bound by the host's ~1600 launches, not by the GPU
        self.wgrad_overlap = os.environ.get("AG3D_WGRAD_STREAM", "0") == "1"
        self.split_rows = True      # keep activations as bf16 hi/lo pair rows between tensor-core layers
        self._fold_cache = None

    # -- per-checkpoint constants (folded BatchNorm scale/shift, bf16 hi/lo weight images for the tensor-core
    #    path), recomputed only when a parameter/buffer changed
    def _apply(self, fn, *a, **kw):
        ops.bump_param_generation()            # .to()/.cuda()/.float() re-seat the parameters
        return super()._apply(fn, *a, **kw)

    def _state_key(self, with_buffers):
        """Cheap change detector for the derived-weight caches: the global parameter generation (raw-kernel writers,
        load_state_dict, .to()) plus torch's own version counters (in-place torch writers, e.g. torch.optim)."""
        ts = self._key_tensors = getattr(self, "_key_tensors", None) or (list(self.parameters()), list(self.buffers()))
        v = sum(t._version for t in ts[0])
        if with_buffers:
            v += sum(t._version for t in ts[1])
        return (self.algo, ops.param_generation(), v, ts[0][0].data_ptr())

    def _folded(self):
        key = self._state_key(True)
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

    def prepare_maps(self, st):
        """coordinate levels, kernel maps and (reorder_rows) the internal row order of a SparseTensor, built once"""
        if st.maps is None:
            st.maps = CoordinateMaps(st.C)
        if self.reorder_rows and st.maps.perm[0] is None and not getattr(st.maps, "reorder_done", False):
            st.maps._reorder()
            st.maps.reorder_done = True
        return st.maps

    @torch.no_grad()
    def forward(self, st):
        """st: agile3d_b200.SparseTensor -> (features [N0, 96], [5 feature maps], CoordinateMaps).  Eval mode only;
        train mode goes through train_forward / train_backward (batch-statistics BatchNorm, saved activations)."""
        if self.training:
            raise RuntimeError("Res16UNet34C.forward is the eval-mode graph; Agile3d.forward_backbone dispatches "
                               "to train_forward in train mode")
        maps, fold, dev = self.prepare_maps(st), self._folded(), st.F.device
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
                          out_split=self.split_rows and self.algo != ops.ALGO_SIMT, bricks=maps.bricks)
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


    # ====================================================================================== train mode
    # Reference semantics: MinkowskiBatchNorm with batch statistics over all voxels of the batch (biased variance
    # for normalisation, unbiased for the running update), autograd through every layer (engine.py:53,146).
    # Here: every conv writes its raw output z, ag3d_bn_stats/ag3d_bn_apply produce y (+residual, ReLU); the
    # backward walks the recorded layers in reverse: ag3d_bn_bwd -> ag3d_spconv_bwd_weight -> ag3d_spconv_bwd_data
    # (the forward kernel over the transposed map, skip-branch gradients added in its epilogue).
    def _train_weights(self):
        """name -> (W [K,cin,cout], W tc image, W_t [K,cout,cin] for the data gradient, W_t tc image)."""
        key = self._state_key(False)
        cache = getattr(self, "_train_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        table = {}
        tc = self.algo != ops.ALGO_SIMT
        for name, mod in self.named_modules():
            if not isinstance(mod, SparseConv) or mod.cin % 32 != 0:
                continue
            w = mod.kernel.detach()
            w3 = (w if w.dim() == 3 else w.unsqueeze(0)).contiguous()
            wt = (w3.flip(0) if w3.shape[0] == 27 else w3).transpose(1, 2).contiguous()
            table[name] = (w3, ops.prepare_tc_weight(w3) if tc else None, wt, self._prep_t(wt) if tc else None)
        self._train_cache = (key, table)
        return table

    @staticmethod
    def _chunks(c):
        """output-channel chunks the tensor-core kernel accepts (<= 256, multiples of 32)."""
        return [(0, c)] if c <= 256 else [(o, min(o + 192, c)) for o in range(0, c, 192)]

    def _prep_t(self, wt):
        return [ops.prepare_tc_weight(wt[:, :, a:b].contiguous()) for a, b in self._chunks(wt.shape[2])]

    def _dgrad(self, name, dz, nbr_t, n_in, W, residual=None, in_split=False):
        """din [n_in, cin] = sum_k dz[nbr_t[k]] @ W_t[k] (+ residual).  in_split: dz holds split (bf16 hi/lo) rows."""
        _, _, wt, wt_tc = W[name]
        cin = wt.shape[2]
        din = torch.empty((n_in, cin), dtype=torch.float32, device=dz.device)
        small = min(n_in, dz.shape[0]) < self.SMALL_LEVEL_ROWS and not in_split
        for i, (a, b) in enumerate(self._chunks(cin)):
            whole = (a, b) == (0, cin)
            ops.spconv_fwd(dz, nbr_t, wt if whole else wt[:, :, a:b].contiguous(), din[:, a:b],
                           residual=None if residual is None else residual[:, a:b],
                           algo=ops.ALGO_SIMT if small else self.algo,
                           weight_tc=wt_tc[i] if (wt_tc is not None and not small) else None, in_split=in_split)
        return din

    # Train mode normalises every conv output with the statistics of THIS batch.  On a level with a handful of rows
    # (the coarsest level of a small scene has ~4) the batch variance of a channel can be orders of magnitude below
    # its mean square, and BatchNorm then amplifies the 1e-5 rounding of the bf16x3 products past the 1e-3 parity
    # criterion.  Such levels carry no measurable time, so they run on the exact-fp32 kernels.
    SMALL_LEVEL_ROWS = 256

    def _tc_rows(self, cin, cout, K=1, rows=None):
        """tensor-core mode and shapes the tcgen05 kernels take: layer inputs and output gradients then travel as
        split (bf16 hi/lo) copies, gathered by the TMA engine in the forward, data-gradient and weight-gradient kernels"""
        if rows is not None and rows < self.SMALL_LEVEL_ROWS:
            return False
        return self.algo != ops.ALGO_SIMT and cin % 32 == 0 and cout % 32 == 0 and ops.wgrad_tc_supported(K, cin, cout)

    def _split_of(self, x):
        """split copy of a layer input, shared by the convolutions that read the same tensor (conv1 and the block's
        downsample; forward and weight gradient)"""
        cache = getattr(self, "_split_cache", None)
        key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()))
        xs = cache.get(key) if cache is not None else None
        if xs is None:
            xs = ops.pack_split_rows(x)
            if cache is not None:
                cache.clear()                      # inputs are consumed block by block: keep only the latest
                cache[key] = xs
        return xs

    def _wgrad_stream(self, dev):
        """side stream of the weight-gradient kernels, owned by the caller's stream (AG3D_WGRAD_STREAM=0: none)"""
        if not self.wgrad_overlap or dev.type != "cuda":
            return None
        cache = self.__dict__.setdefault("_wg_streams", {})
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        if key not in cache:
            cache[key] = torch.cuda.Stream(device=dev)
        return cache[key]

    def _wgrad_join(self, dev):
        side = self._wgrad_stream(dev)
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)

    def _wgrad(self, x, nbr, dy, K, xs=None, dys=None):
        """dW = x[nbr]^T dy: tcgen05 kernel on split copies of both operands in tensor-core mode, else fp32 SIMT."""
        cin, cout = x.shape[1], dy.shape[1]
        if not self._tc_rows(cin, cout, K, min(x.shape[0], dy.shape[0])):
            return ops.spconv_bwd_weight(x, nbr, dy, K)
        return ops.spconv_bwd_weight_tc(xs if xs is not None else self._split_of(x), nbr,
                                        dys if dys is not None else ops.pack_split_rows(dy), K)

    def _conv_bn_fwd(self, name, bn_name, x, nbr, nbr_t, n_out, out, W, residual=None, relu=True):
        w3, wtc = W[name][0], W[name][1]
        bn = self.get_submodule(bn_name).bn
        z = torch.empty((n_out, w3.shape[2]), dtype=torch.float32, device=x.device)
        small = min(n_out, x.shape[0]) < self.SMALL_LEVEL_ROWS
        xs = self._split_of(x) if (wtc is not None and self._tc_rows(w3.shape[1], w3.shape[2], w3.shape[0],
                                                                     min(n_out, x.shape[0]))) else None
        ops.spconv_fwd(x if xs is None else xs, nbr, w3, z, algo=ops.ALGO_SIMT if small else self.algo,
                       weight_tc=None if small else wtc, in_split=xs is not None)
        mean, invstd = ops.bn_stats(z, bn.eps, bn.momentum, bn.running_mean, bn.running_var)
        bn.num_batches_tracked += 1
        ops.bn_apply(z, mean, invstd, bn.weight.detach(), bn.bias.detach(), out, residual=residual, relu=relu)
        return dict(name=name, bn=bn_name, x=x, xs=xs, nbr=nbr, nbr_t=nbr_t, z=z, y=out, mean=mean, invstd=invstd, relu=relu)

    def _conv_bn_bwd(self, rec, dy, W, grads, g_out=None, want_dx=True, dx_residual=None):
        """dy: gradient of the layer's output y (overwritten with dz).  -> din (or None)."""
        bn = self.get_submodule(rec["bn"]).bn
        dgamma, dbeta = ops.bn_bwd(rec["z"], rec["y"] if rec["relu"] else None, dy, rec["mean"], rec["invstd"],
                                   bn.weight.detach(), dy, relu=rec["relu"], g_out=g_out)
        grads[rec["bn"] + ".bn.weight"] = dgamma
        grads[rec["bn"] + ".bn.bias"] = dbeta
        K = W[rec["name"]][0].shape[0]
        dys = ops.pack_split_rows(dy) if self._tc_rows(rec["x"].shape[1], dy.shape[1], K,
                                                       min(rec["x"].shape[0], dy.shape[0])) else None
        # the weight gradient is a leaf of the backward: it runs on a side stream under the data-gradient chain
        side = self._wgrad_stream(dy.device)
        if side is not None:
            xs = rec.get("xs")
            if xs is None and dys is not None:
                xs = self._split_of(rec["x"])              # split copy made on the main stream, before the fork
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            side.wait_event(ev)
            with torch.cuda.stream(side):
                dw = self._wgrad(rec["x"], rec["nbr"], dy, K, xs=xs, dys=dys)
            for t in (rec["x"], xs, dy, dys, rec["nbr"]):
                if t is not None:
                    t.record_stream(side)                  # keep the allocator from recycling them under the side stream
        else:
            dw = self._wgrad(rec["x"], rec["nbr"], dy, K, xs=rec.get("xs"), dys=dys)
        kernel = self.get_submodule(rec["name"]).kernel
        grads[rec["name"] + ".kernel"] = dw.view_as(kernel)
        if not want_dx:
            return None
        return self._dgrad(rec["name"], dy if dys is None else dys, rec["nbr_t"], rec["x"].shape[0], W,
                           residual=dx_residual, in_split=dys is not None)

    def _block_train_fwd(self, tape, prefix, blk, x, nbr, W, out=None):
        n, planes = x.shape[0], blk.conv1.cout
        f32 = dict(dtype=torch.float32, device=x.device)
        t = torch.empty((n, planes), **f32)
        r1 = self._conv_bn_fwd(prefix + ".conv1", prefix + ".norm1", x, nbr, nbr, n, t, W, relu=True)
        rd = None
        if blk.downsample is not None:
            res = torch.empty((n, planes), **f32)
            rd = self._conv_bn_fwd(prefix + ".downsample.0", prefix + ".downsample.1", x, None, None, n, res, W,
                                   relu=False)
        else:
            res = x
        if out is None:
            out = torch.empty((n, planes), **f32)
        r2 = self._conv_bn_fwd(prefix + ".conv2", prefix + ".norm2", t, nbr, nbr, n, out, W, residual=res, relu=True)
        tape.append(("block", (r1, r2, rd)))
        return out

    def _block_train_bwd(self, blk_rec, dout, W, grads):
        r1, r2, rd = blk_rec
        g2 = torch.empty_like(dout)                       # relu-masked dout = gradient of the residual operand
        dt = self._conv_bn_bwd(r2, dout, W, grads, g_out=g2)
        if rd is not None:
            dxr = self._conv_bn_bwd(rd, g2, W, grads)
        else:
            dxr = g2
        return self._conv_bn_bwd(r1, dt, W, grads, dx_residual=dxr)

    @torch.no_grad()
    def train_forward(self, st):
        """-> (features [N0,96], [5 feature maps], maps, tape).  tape feeds train_backward."""
        maps, W, dev = self.prepare_maps(st), self._train_weights(), st.F.device
        N, P = maps.sizes, PLANES
        f32 = dict(dtype=torch.float32, device=dev)
        skip_c = (INIT_DIM, P[0], P[1], P[2])
        up_c = (P[7], P[6], P[5], P[4])
        cat = [torch.empty((N[l], up_c[l] + skip_c[l]), **f32) for l in range(4)]
        tape = []
        self._split_cache = {}
        # stem: raw conv -> bn0 (batch statistics) -> relu
        bn0 = self.bn0.bn
        z0 = torch.empty((N[0], INIT_DIM), **f32)
        ops.stem_conv_fwd(maps.coords[0], st.F, maps.tables[0], maps.caps[0], self.conv1_kernel_size,
                          self.conv0p1s1.kernel.detach(), z0, None, None, relu=False, bricks=maps.bricks)
        mean, invstd = ops.bn_stats(z0, bn0.eps, bn0.momentum, bn0.running_mean, bn0.running_var)
        bn0.num_batches_tracked += 1
        y = cat[0][:, up_c[0]:]
        ops.bn_apply(z0, mean, invstd, bn0.weight.detach(), bn0.bias.detach(), y, relu=True)
        stem = dict(z=z0, y=y, mean=mean, invstd=invstd, feats=st.F)
        for i, tag in enumerate(_ENC):
            d = torch.empty((N[i + 1], getattr(self, f"conv{tag}s2").cout), **f32)
            tape.append(("conv", self._conv_bn_fwd(f"conv{tag}s2", f"bn{i + 1}", y, maps.down[i], maps.up[i], N[i + 1],
                                                   d, W)))
            blocks = getattr(self, f"block{i + 1}")
            for b, blk in enumerate(blocks):
                last = b == len(blocks) - 1
                d = self._block_train_fwd(tape, f"block{i + 1}.{b}", blk, d, maps.k3[i + 1], W,
                                          out=cat[i + 1][:, up_c[i + 1]:] if (last and i < 3) else None)
            y = d
        fmaps = [y]
        for j, tag in enumerate(_DEC):
            lvl = 3 - j
            tape.append(("conv", self._conv_bn_fwd(f"convtr{tag}s2", f"bntr{4 + j}", y, maps.up[lvl], maps.down[lvl],
                                                   N[lvl], cat[lvl][:, :up_c[lvl]], W)))
            y = cat[lvl]
            for b, blk in enumerate(getattr(self, f"block{5 + j}")):
                y = self._block_train_fwd(tape, f"block{5 + j}.{b}", blk, y, maps.k3[lvl], W)
            fmaps.append(y)
        self._split_cache = None
        return y, fmaps, maps, (tape, stem, up_c)

    @torch.no_grad()
    def train_backward(self, maps, saved, dy, sink=None):
        """dy: gradient of the backbone output [N0,96] (consumed).  -> {parameter name: gradient}.
        sink (optional): called with the gradients of every finished stage, last stage of the forward first, as soon as
        they exist (the data-parallel exchange of that stage then overlaps the rest of the backward, optim.GradBuckets);
        gradients handed to the sink are not returned."""
        tape, stem, up_c = saved
        W = self._train_weights()
        self._split_cache = {}
        grads = {}
        tape = list(tape)
        skip_grads = {}

        def pop(kind):
            k, rec = tape.pop()
            assert k == kind, "tape walk out of step with the graph"
            return rec

        def stage_done():
            self._wgrad_join(dy.device)                            # the stage's weight gradients are complete
            if sink is not None and grads:
                sink(dict(grads))
                grads.clear()

        for j in reversed(range(4)):                               # decoder stages, level 0 first
            lvl = 3 - j
            for _ in range(len(getattr(self, f"block{5 + j}"))):
                dy = self._block_train_bwd(pop("block"), dy, W, grads)
            skip_grads[lvl] = dy[:, up_c[lvl]:]                    # gradient of the skip half of the concat
            dy = self._conv_bn_bwd(pop("conv"), dy[:, :up_c[lvl]], W, grads)       # transposed conv -> coarser level
            stage_done()
        for e in reversed(range(4)):                               # encoder stages, deepest first
            for _ in range(len(getattr(self, f"block{e + 1}"))):
                dy = self._block_train_bwd(pop("block"), dy, W, grads)
            # the stride-2 conv's input (level e) also feeds the level-e concat: add that gradient in the epilogue
            dy = self._conv_bn_bwd(pop("conv"), dy, W, grads, dx_residual=skip_grads[e])
            stage_done()
        assert not tape, "tape walk out of step with the graph"
        dp1 = dy                                                   # total gradient of the stem output
        bn0 = self.bn0.bn
        dgamma, dbeta = ops.bn_bwd(stem["z"], stem["y"], dp1, stem["mean"], stem["invstd"], bn0.weight.detach(), dp1,
                                   relu=True)
        grads["bn0.bn.weight"], grads["bn0.bn.bias"] = dgamma, dbeta
        self._split_cache = None
        grads["conv0p1s1.kernel"] = ops.stem_bwd_weight(maps.coords[0], stem["feats"], maps.tables[0], maps.caps[0],
                                                        self.conv1_kernel_size, dp1,
                                                        bricks=maps.bricks).view_as(self.conv0p1s1.kernel)
        stage_done()
        return grads
