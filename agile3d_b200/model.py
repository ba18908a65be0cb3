"""Agile3d: the reference's model surface (models/agile3d.py:19-421) on the B200 library.

    model = build_agile3d(args)                       # models/agile3d.py:399-421
    pcd_features, aux, coordinates, pos_encodings_pcd = model.forward_backbone(x, raw_coordinates)   # :163-181
    out = model.forward_mask(pcd_features, aux, coordinates, pos_encodings_pcd, click_idx, click_time_idx)  # :183-339
    out['pred_masks'][b]  -> [Nv_b, 1 + K_b] fp32 logits, column index = object id

The four values returned by ``forward_backbone`` are opaque handles, exactly as every reference caller treats
them.  state_dict keys/shapes equal the reference's.

Per click round the voxel features are streamed by exactly two kernels per decoder layer
(ag3d_c2s_attn_fwd, ag3d_s2c_mask_fwd); the O(Nq) glue between them (projections of the <= few dozen click
queries, click<->click self-attention, FFN) is tiny and stays in torch.
"""
from __future__ import annotations

import math

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .backbone import PLANES, Res16UNet34C, SparseConv
from .minkowski import SparseTensor


class _AttnParams(nn.Module):
    """Parameter layout of nn.MultiheadAttention (in_proj_weight [3d,d], in_proj_bias, out_proj.{weight,bias})."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.zeros_(self.out_proj.bias)


class _CrossLayer(nn.Module):           # models/modules/attention_block.py:64-98
    def __init__(self, d):
        super().__init__()
        self.multihead_attn = _AttnParams(d)
        self.norm = nn.LayerNorm(d)


class _SelfLayer(nn.Module):            # models/modules/attention_block.py:5-38
    def __init__(self, d):
        super().__init__()
        self.self_attn = _AttnParams(d)
        self.norm = nn.LayerNorm(d)


class _FFNLayer(nn.Module):             # models/modules/attention_block.py:127-155
    def __init__(self, d, dff):
        super().__init__()
        self.linear1 = nn.Linear(d, dff)
        self.linear2 = nn.Linear(dff, d)
        self.norm = nn.LayerNorm(d)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)


class _FourierPos(nn.Module):           # models/position_embedding.py:44-72 (buffer is part of the checkpoint)
    def __init__(self, d, gauss_scale):
        super().__init__()
        self.register_buffer("gauss_B", torch.empty((3, d // 2)).normal_() * gauss_scale)


def _time_table(d_model, length):       # models/position_embedding.py:210-226
    pe = torch.zeros(length, d_model)
    pos = torch.arange(0, length).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class BackboneFeatures:
    """Opaque handle for ``pcd_features`` / ``coordinates``: features of all scenes + per-scene row ranges.

    The backbone works in an INTERNAL row order (rows of a scene sorted by neighbour pattern, CoordinateMaps._reorder);
    ``Fp`` holds the rows in that order, ``perm[new] = caller row`` and ``inv[caller row] = new``.  ``F`` is the caller's
    view (gathered on first use); scenes keep their row ranges in both orders.

    In tensor-core eval mode the features exist as ``Fs``: "split" rows (32-channel slabs of bf16 hi | bf16 lo, the
    backbone's activation format), which is what the TMA-fed decoder kernels stream; the fp32 rows ``Fp`` are then
    decoded on first use (callers that read ``.F``, click rounds with more than 24 queries)."""

    def __init__(self, F_, offsets, C=None, perm=None, inv=None, Fs=None):
        self._Fp, self.Fs, self.offsets, self.C, self.perm, self.inv = F_, Fs, offsets, C, perm, inv
        self._F, self._inv_local, self._inv64, self._inv_local32 = None, None, None, None

    @property
    def Fp(self):
        if self._Fp is None:
            self._Fp = ops.unpack_split(self.Fs)
        return self._Fp

    @property
    def F(self):
        if self.inv is None:
            return self.Fp
        if self._F is None:
            self._F = self.Fp.index_select(0, self.inv64)
        return self._F

    @property
    def inv64(self):
        if self._inv64 is None and self.inv is not None:
            self._inv64 = self.inv.long()
        return self._inv64

    def rows_internal(self, rows):
        """caller row indices (device int64 tensor) -> internal rows"""
        return rows if self.inv is None else self.inv64[rows]

    def to_caller(self, b, t):
        """rows of scene b in internal order -> caller order (differentiable)"""
        if self.inv is None:
            return t
        if not t.requires_grad and t.dim() == 2 and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
            if self._inv_local32 is None:
                self._inv_local32 = [self.inv[self.offsets[i]:self.offsets[i + 1]] - self.offsets[i]
                                     for i in range(len(self.offsets) - 1)]
            return ops.gather_rows(t, self._inv_local32[b])
        if self._inv_local is None:
            self._inv_local = [self.inv64[self.offsets[i]:self.offsets[i + 1]] - self.offsets[i]
                               for i in range(len(self.offsets) - 1)]
        return t.index_select(0, self._inv_local[b])

    @property
    def decomposed_features(self):
        return [self.F[self.offsets[b]:self.offsets[b + 1]] for b in range(len(self.offsets) - 1)]

    @property
    def device(self):
        return (self.Fs if self._Fp is None else self._Fp).device


class _AuxMap:
    """One of the 5 backbone feature maps returned as ``aux``: rows in the backbone's internal order; ``dense()`` decodes
    the bf16 hi/lo pair rows of the tensor-core eval mode into fp32."""

    def __init__(self, rows, split):
        self.rows, self.split = rows, split
        self.shape = rows.shape

    def dense(self):
        return ops.unpack_split(self.rows) if self.split else self.rows


class _PosList:
    """per-scene positional encodings: ``[b]`` gives the caller's row order (as the reference's list would), the model
    itself reads ``internal[b]`` (the backbone's row order)."""

    def __init__(self, internal, handle, split=None):
        self.internal, self._h, self.split = internal, handle, split      # split: the same rows as bf16 hi/lo pairs

    def __len__(self):
        return len(self.internal)

    def __getitem__(self, b):
        return self._h.to_caller(b, self.internal[b])

    def __iter__(self):
        return (self[b] for b in range(len(self.internal)))


# ================================================================================================ autograd bridges
# torch.autograd is the plumbing that routes gradients between the query-side glue (tiny torch ops) and the CUDA
# kernels; every Function below is a pair of calls into the library (forward kernel, backward kernel).
class _BackboneFn(torch.autograd.Function):
    """Res16UNet34C + lin_squeeze_head in train mode: one node whose backward is Res16UNet34C.train_backward."""

    @staticmethod
    def forward(ctx, model, st, names, holder, *params):
        bb, head = model.backbone, model.lin_squeeze_head
        feats, fmaps, maps, saved = bb.train_forward(st)
        holder["fmaps"] = fmaps
        tc = bb.algo != ops.ALGO_SIMT
        w = head.kernel.detach()
        pcd = torch.empty((feats.shape[0], head.cout), dtype=torch.float32, device=feats.device)
        ops.spconv_fwd(feats, None, w, pcd, None, head.bias.detach().reshape(-1).contiguous(), algo=bb.algo,
                       weight_tc=ops.prepare_tc_weight(w) if tc else None)
        ctx.model, ctx.maps, ctx.saved, ctx.feats, ctx.names = model, maps, saved, feats, names
        return pcd

    @staticmethod
    def backward(ctx, dpcd):
        model, feats = ctx.model, ctx.feats
        bb, head = model.backbone, model.lin_squeeze_head
        dpcd = dpcd.contiguous()
        sink = getattr(model, "grad_sink", None)          # optim.GradBuckets: gradients go straight to the flat buffer
        with torch.no_grad():
            grads = {"lin_squeeze_head.kernel": bb._wgrad(feats, None, dpcd, 1).view_as(head.kernel),
                     "lin_squeeze_head.bias": ops.col_sum(dpcd).view_as(head.bias)}
            if sink is not None:
                # every gradient outside the backbone is final now: autograd accumulates leaf gradients as soon as they
                # are ready (AccumulateGrad nodes run first), and this node is the last consumer of the decoder's inputs
                sink.write(grads)
                sink.tail_done()
                grads = {}
            wt = head.kernel.detach().t().contiguous()
            dfeats = torch.empty_like(feats)
            ops.spconv_fwd(dpcd, None, wt, dfeats, algo=bb.algo,
                           weight_tc=ops.prepare_tc_weight(wt) if bb.algo != ops.ALGO_SIMT else None)
            cb = None if sink is None else (lambda g: sink.stage({"backbone." + k: v for k, v in g.items()}))
            for k, v in bb.train_backward(ctx.maps, ctx.saved, dfeats, sink=cb).items():
                grads["backbone." + k] = v
        ctx.saved = None
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


# X^T dY contractions over the voxels of the decoder backward (query-side gradients): tcgen05 weight-gradient kernel on
# split copies of both operands when the model runs in tensor-core mode (set by Agile3d._forward_mask), else fp32 SIMT
_WGRAD_TC = [False]


def _xt_dy(xs, dys, tc):
    """sum over xs of x^T dy -> [1, cin, cout]; xs is a list of [Nv, cin] operands sharing one dy [Nv, cout]."""
    dy = dys
    cin, cout = xs[0].shape[1], dy.shape[1]
    if tc and ops.wgrad_tc_supported(1, cin, cout):
        dsplit = ops.pack_split_rows(dy)
        out = None
        for x in xs:
            out = ops.spconv_bwd_weight_tc(ops.pack_split_rows(x), None, dsplit, 1, dweight=out, accumulate=out is not None)
        return out
    out = None
    for x in xs:
        out = ops.spconv_bwd_weight(x, None, dy, 1, dweight=out, accumulate=out is not None)
    return out


def _pad_rows(t, rows):
    out = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


class _C2sFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, qfold, nq, H, label, q_obj, obj_count):
        out = torch.empty((H * nq, x.shape[1]), dtype=torch.float32, device=x.device)
        lse = torch.empty(H * nq, dtype=torch.float32, device=x.device)
        qfold = qfold.contiguous()
        ops.c2s_attn_fwd(x, pos, qfold, nq, H, label, q_obj, obj_count, out=out, lse=lse)
        ctx.save_for_backward(x, pos, qfold, out, lse)
        ctx.meta = (nq, H, label, q_obj, obj_count)
        ctx.tc = _WGRAD_TC[0]                      # the mode of the model that ran this forward
        return out

    @staticmethod
    def backward(ctx, dctx):
        x, pos, qfold, out, lse = ctx.saved_tensors
        nq, H, label, q_obj, obj_count = ctx.meta
        HQ = H * nq
        # tensor-core mode takes any number of queries (rows padded to a multiple of 32, processed in chunks of 256);
        # the exact-fp32 kernels hold all (head, query) rows of a voxel tile at once: at most 32 queries
        hqp = ops._pad32(HQ) if (ctx.tc and HQ > 256) else ops.decoder_bwd_rows(nq, H)
        with torch.no_grad():
            dctx = dctx.contiguous()
            qf, dc = _pad_rows(qfold, hqp), _pad_rows(dctx, hqp)
            rowobj = torch.full((hqp,), -2, dtype=torch.int32, device=x.device)
            if label is None:
                ro = torch.full((HQ,), -1, dtype=torch.int32, device=x.device)
            else:
                o = q_obj.long()
                ro = torch.where(obj_count[o] > 0, q_obj, torch.full_like(q_obj, -1)).repeat(H)     # rows are h-major
            rowobj[:HQ] = torch.where(torch.isinf(lse), torch.full_like(ro, -2), ro)
            lse_p = torch.full((hqp,), float("inf"), dtype=torch.float32, device=x.device)
            lse_p[:HQ] = lse
            dr = _pad_rows((dctx * out).sum(1), hqp)
            if ctx.tc:            # tensor-core mode: the four GEMMs as 1x1 tcgen05 convolutions
                dx, ds_chunks = ops.c2s_attn_bwd_tc(x, pos, qf, dc, lse_p, dr, rowobj, hqp, label)
                xp = x + pos
                dq = torch.cat([_xt_dy([ds], xp, True)[0] for ds in ds_chunks], 0)       # dS^T x + dS^T pos
            else:
                dx, ds = ops.c2s_attn_bwd(x, pos, qf, qf.t().contiguous(), dc, dc.t().contiguous(), lse_p, dr, rowobj,
                                          hqp, label)
                dq = _xt_dy([ds], x + pos, False)[0]
        return dx, None, dq[:HQ], None, None, None, None, None


class _S2cFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, A, c, U, bo, ln_w, ln_b, E, eps, q_obj, nq, H, n_obj):
        A, c, U, E = A.contiguous(), c.contiguous(), U.contiguous(), E.contiguous()
        x_out, logits, label, count = ops.s2c_mask_fwd(x, pos, A, c, U, bo, ln_w, ln_b, eps, E, q_obj, nq, H, n_obj)
        ctx.save_for_backward(x, pos, A, c, U, bo, ln_w, ln_b, E, x_out)
        ctx.meta = (eps, q_obj, nq, H, n_obj)
        ctx.tc = _WGRAD_TC[0]                      # the mode of the model that ran this forward
        ctx.mark_non_differentiable(label, count)
        ctx.set_materialize_grads(False)
        return x_out, logits, label, count

    @staticmethod
    def backward(ctx, dxo, dlogits, _dl, _dc):
        x, pos, A, c, U, bo, ln_w, ln_b, E, x_out = ctx.saved_tensors
        eps, q_obj, nq, H, n_obj = ctx.meta
        HQ = H * nq
        with torch.no_grad():
            dxo_c = None if dxo is None else dxo.contiguous()
            dlg_c = None if dlogits is None else dlogits.contiguous()
            if ctx.tc:            # tensor-core mode: GEMMs as 1x1 tcgen05 convolutions, any number of queries
                dx, dA, dc, dU, dbo, dlw, dlb, dE = ops.s2c_mask_bwd_tc_any(
                    x, pos, A, c, U, bo, ln_w, ln_b, eps, E, q_obj, nq, H, n_obj, dxo_c, dlg_c, x_out,
                    lambda xs, dy: _xt_dy(xs, dy, True))
                return dx, None, dA, dc, dU, dbo, dlw, dlb, dE, None, None, None, None, None
            hqp = ops.decoder_bwd_rows(nq, H)
            Ap, Up, Ep = _pad_rows(A, hqp), _pad_rows(U, hqp), _pad_rows(E, 32)
            dx, a, ds, dy, g, cols = ops.s2c_mask_bwd(
                x, pos, Ap, Ap.t().contiguous(), _pad_rows(c, hqp), Up, Up.t().contiguous(), bo, ln_w, ln_b, eps, Ep,
                Ep.t().contiguous(), q_obj, nq, H, n_obj, hqp, dxo_c, dlg_c)
            dA = _xt_dy([ds], x + pos, False)               # dS^T x + dS^T pos
            dU = _xt_dy([a], dy, False)
            dE = _xt_dy([g], x_out, False)
        return (dx, None, dA[0, :HQ], cols[384:384 + HQ], dU[0, :HQ], cols[:128], cols[128:256], cols[256:384],
                dE[0, :nq], None, None, None, None, None)


class Agile3d(nn.Module):
    def __init__(self, backbone, hidden_dim, num_heads, dim_feedforward, shared_decoder, num_decoders,
                 num_bg_queries, dropout, pre_norm, positional_encoding_type, normalize_pos_enc, hlevels,
                 voxel_size, gauss_scale, aux):
        super().__init__()
        if hidden_dim != 128 or num_heads != 8:
            raise ValueError("the decoder kernels are specialised for hidden_dim=128, num_heads=8 (reference defaults)")
        if positional_encoding_type != "fourier" or not normalize_pos_enc:
            raise ValueError("only the reference default positional encoding (fourier, normalised) is built")
        if pre_norm or dropout != 0.0:
            raise ValueError("only post-norm, dropout=0 (reference defaults) are built")
        if list(hlevels) != [4]:
            raise ValueError("only hlevels=[4] (reference default: full-resolution voxels) is built")
        self.hidden_dim, self.num_heads, self.num_decoders = hidden_dim, num_heads, num_decoders
        self.num_bg_queries, self.shared_decoder, self.aux = num_bg_queries, shared_decoder, aux
        self.hlevels, self.voxel_size = list(hlevels), voxel_size
        d = hidden_dim
        self.backbone = backbone
        self.lin_squeeze_head = SparseConv(PLANES[7], d, 1, bias=True)
        self.bg_query_feat = nn.Embedding(num_bg_queries, d)
        self.bg_query_pos = nn.Embedding(num_bg_queries, d)
        self.mask_embed_head = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, d))
        self.pos_enc = _FourierPos(d, gauss_scale)
        n_shared = 1 if shared_decoder else num_decoders

        def stack(make):
            return nn.ModuleList([nn.ModuleList([make() for _ in self.hlevels]) for _ in range(n_shared)])

        self.c2s_attention = stack(lambda: _CrossLayer(d))
        self.s2c_attention = stack(lambda: _CrossLayer(d))
        self.c2c_attention = stack(lambda: _SelfLayer(d))
        self.ffn_attention = stack(lambda: _FFNLayer(d, dim_feedforward))
        self.decoder_norm = nn.LayerNorm(d)
        self.time_encode = _time_table(d, 200)       # plain attribute, not in the state_dict (agile3d.py:138)
        self.fused_queries = True                    # eval: click-query side in fused kernels (csrc/query_ops.cu)
        self.split_decoder = True                    # eval, tensor-core mode: voxel features / encodings as split rows (TMA-fed decoder)
        self.decoder_streams = int(os.environ.get("AG3D_DEC_STREAMS", "4"))   # eval: the per-scene voxel kernels of a layer alternate between a few side
        #                                              streams, so the prologue / tail / merge of one scene run under the next
        # derived weight images (folded BatchNorm, tensor-core images) are cached per parameter generation
        self.register_load_state_dict_post_hook(lambda module, incompatible: ops.bump_param_generation())

    def _apply(self, fn, *a, **kw):
        ops.bump_param_generation()
        return super()._apply(fn, *a, **kw)

    # ------------------------------------------------------------------------------------------ backbone
    def forward_backbone(self, x: SparseTensor, raw_coordinates=None):
        if not isinstance(x, SparseTensor):
            raise TypeError("forward_backbone expects an agile3d_b200.SparseTensor")
        raw = raw_coordinates.to(device=x.F.device, dtype=torch.float32).contiguous()
        if raw.shape != (x.F.shape[0], 3):
            raise ValueError("raw_coordinates must be [N,3]")
        # scene row ranges (scenes are contiguous and ordered: SURVEY.md A.2); coordinate maps + the internal row order
        offsets = x.scene_offsets()
        n_scenes = len(offsets) - 1
        with torch.no_grad():
            maps = self.backbone.prepare_maps(x)
            perm, inv = maps.perm[0], maps.inv[0]
            raw_i = raw if perm is None else ops.gather_rows(raw, perm)                # xyz in the internal row order
            split_dec = (not self.training) and self.split_decoder and self.backbone.split_rows \
                and self.backbone.algo != ops.ALGO_SIMT and self.hidden_dim == 128
            pos_s = None
            # the encodings do not depend on the backbone: they are computed on a side stream under its kernels (eval)
            aux = self._side_streams(raw_i.device, 1) if (not self.training and raw_i.is_cuda and self.decoder_streams) else None
            self._fork(aux)
            with self._on(aux, 0):
                if split_dec:
                    pos, rng, pos_s = ops.fourier_posenc(raw_i, offsets, self.pos_enc.gauss_B, want_split=True)
                else:
                    pos, rng = ops.fourier_posenc(raw_i, offsets, self.pos_enc.gauss_B)
        if self.training:
            # batch-statistics BatchNorm + recorded activations; one autograd node for backbone + head (engine.py:53)
            named = [(n, p) for n, p in self.named_parameters()
                     if n.startswith("backbone.") or n.startswith("lin_squeeze_head.")]
            holder = {}
            pcd = _BackboneFn.apply(self, x, tuple(n for n, _ in named), holder, *[p for _, p in named])
            fmaps = holder["fmaps"]
        else:
            with torch.no_grad():
                pcd, fmaps = self._forward_backbone_eval(x, split_dec)
        self._join(aux)
        pcd_features = BackboneFeatures(None, offsets, x.C, perm, inv, Fs=pcd) if split_dec \
            else BackboneFeatures(pcd, offsets, x.C, perm, inv)
        coordinates = BackboneFeatures(raw, offsets, x.C)
        coordinates.range = rng
        # only the full-resolution level is ever read (hlevels=[4], agile3d.py:278); keep the reference's indexing
        plist = _PosList([pos[offsets[b]:offsets[b + 1]] for b in range(n_scenes)], pcd_features,
                         None if pos_s is None else [pos_s[offsets[b]:offsets[b + 1]] for b in range(n_scenes)])
        pos_encodings_pcd = [None, None, None, None, [plist]]
        return pcd_features, fmaps, coordinates, pos_encodings_pcd

    def _forward_backbone_eval(self, x, out_split=False):
        feats, fmaps, maps = self.backbone(x)
        pcd = torch.empty((feats.shape[0], self.hidden_dim), dtype=torch.float32, device=feats.device)
        head = self.lin_squeeze_head
        hkey = (self.backbone.algo, ops.param_generation(), head.kernel.data_ptr(), head.kernel._version)
        if getattr(self, "_head_tc", (None, None))[0] != hkey:
            wtc = ops.prepare_tc_weight(head.kernel) if self.backbone.algo != ops.ALGO_SIMT else None
            self._head_tc = (hkey, wtc)
        split = self.backbone.split_rows and self.backbone.algo != ops.ALGO_SIMT
        ops.spconv_fwd(feats, None, head.kernel, pcd, None, head.bias.detach().reshape(-1).contiguous(), relu=False,
                       algo=self.backbone.algo, weight_tc=self._head_tc[1], in_split=split, out_split=out_split)
        # the 5 feature maps (`aux`) are opaque to every caller of the reference; in split mode they hold bf16 hi/lo pair
        # rows in the internal row order - hand out objects that say so instead of tensors that look like features
        return pcd, [_AuxMap(f, split) for f in fmaps]

    # ------------------------------------------------------------------------------------------ decoder glue
    # The O(Nq) query-side algebra is batched over all scenes of a batch that have the same number of queries
    # (leading dimension B), so its launch count does not grow with the batch size.
    def _click_pos(self, xyz, lo, hi):
        """fourier encoding of click coordinates (position_embedding.py:123-152): xyz, lo, hi [n,3] -> [n,128]."""
        u = (xyz - lo) / (hi - lo)
        t = (u * (2 * math.pi)) @ self.pos_enc.gauss_B
        return torch.cat([t.sin(), t.cos()], dim=1)

    @staticmethod
    def _fold_c2s(p, tgt, qpos, H):
        """qfold[b,(h,q),:] = Wk_h^T ((Wq_h (tgt+qpos) + bq_h) / sqrt(dh));  tgt, qpos [B,Q,d] -> [B,H*Q,d]."""
        B, Q, d = tgt.shape
        dh = d // H
        Wq, Wk = p.in_proj_weight[:d], p.in_proj_weight[d:2 * d]
        qp = F.linear(tgt + qpos, Wq, p.in_proj_bias[:d]).view(B, Q, H, dh) * (1.0 / math.sqrt(dh))
        return torch.einsum("bqhd,hdc->bhqc", qp, Wk.view(H, dh, d)).reshape(B, H * Q, d).contiguous()

    @staticmethod
    def _finish_c2s(layer, tgt, ctx, H):
        p = layer.multihead_attn
        B, Q, d = tgt.shape
        dh = d // H
        Wv, bv = p.in_proj_weight[2 * d:], p.in_proj_bias[2 * d:]
        heads = torch.einsum("bhqc,hdc->bqhd", ctx.view(B, H, Q, d), Wv.view(H, dh, d)) + bv.view(H, dh)
        attn = F.linear(heads.reshape(B, Q, d), p.out_proj.weight, p.out_proj.bias)
        return layer.norm(tgt + attn)

    @staticmethod
    def _self_attn(layer, tgt, qpos, H):
        p = layer.self_attn
        B, Q, d = tgt.shape
        dh = d // H
        qk = F.linear(tgt + qpos, p.in_proj_weight[:2 * d], p.in_proj_bias[:2 * d])
        v = F.linear(tgt, p.in_proj_weight[2 * d:], p.in_proj_bias[2 * d:]).view(B, Q, H, dh).transpose(1, 2)
        q = qk[..., :d].reshape(B, Q, H, dh).transpose(1, 2)
        k = qk[..., d:].reshape(B, Q, H, dh).transpose(1, 2)
        a = torch.softmax((q * (1.0 / math.sqrt(dh))) @ k.transpose(2, 3), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, Q, d)
        return layer.norm(tgt + F.linear(o, p.out_proj.weight, p.out_proj.bias))

    @staticmethod
    def _ffn(layer, tgt):
        return layer.norm(tgt + layer.linear2(F.relu(layer.linear1(tgt))))

    @staticmethod
    def _fold_s2c(p, queries, qpos, H):
        """A[b,(h,q),:] = Wq_h^T k_hq / sqrt(dh); c[b,(h,q)] = bq_h . k_hq / sqrt(dh); U[b,(h,q),:] = Wo[:,h] v_hq."""
        B, Q, d = queries.shape
        dh = d // H
        Wq, bq = p.in_proj_weight[:d], p.in_proj_bias[:d]
        kp = F.linear(queries + qpos, p.in_proj_weight[d:2 * d], p.in_proj_bias[d:2 * d]).view(B, Q, H, dh)
        vp = F.linear(queries, p.in_proj_weight[2 * d:], p.in_proj_bias[2 * d:]).view(B, Q, H, dh)
        sc = 1.0 / math.sqrt(dh)
        A = (torch.einsum("bqhd,hdc->bhqc", kp, Wq.view(H, dh, d)) * sc).reshape(B, H * Q, d).contiguous()
        c = (torch.einsum("bqhd,hd->bhq", kp, bq.view(H, dh)) * sc).reshape(B, H * Q).contiguous()
        U = torch.einsum("chd,bqhd->bhqc", p.out_proj.weight.view(d, H, dh), vp).reshape(B, H * Q, d).contiguous()
        return A, c, U

    # ------------------------------------------------------------------------------------------ forward_mask
    def forward_mask(self, pcd_features, aux, coordinates, pos_encodings_pcd, click_idx=None, click_time_idx=None):
        """Same kernels with or without autograd: in train mode (engine.py:119-121) the two voxel-streaming kernels run as
        autograd nodes (their backward is ag3d_c2s_attn_bwd / ag3d_s2c_mask_bwd) and the voxel features of every
        layer are kept; otherwise layers > 0 update the features in place."""
        grad = torch.is_grad_enabled() and (self.training or (pcd_features.Fs is None and pcd_features.Fp.requires_grad))
        if not grad:
            with torch.no_grad():
                if self.fused_queries:
                    return self._forward_mask_fused(pcd_features, coordinates, pos_encodings_pcd, click_idx, click_time_idx)
                return self._forward_mask(pcd_features, coordinates, pos_encodings_pcd, click_idx, click_time_idx, False)
        return self._forward_mask(pcd_features, coordinates, pos_encodings_pcd, click_idx, click_time_idx, True)

    # ---- eval path: the click-query side of every layer in three fused kernels (csrc/query_ops.cu) - no torch op
    #      touches the data between the C-ABI calls; train mode keeps the torch glue below for autograd
    def _layer_blob(self, li):
        """per-layer weight blob of ag3d_query_* (pre-transposed matrices, layout in csrc/query_ops.cu)"""
        key = (ops.param_generation(), sum(p._version for p in self._dec_params()))
        cache = getattr(self, "_blob_cache", None)
        if cache is None or cache[0] != key:
            cache = self._blob_cache = (key, {})
        if li not in cache[1]:
            d = self.hidden_dim
            c2s, c2c = self.c2s_attention[li][0], self.c2c_attention[li][0]
            ffn, s2c = self.ffn_attention[li][0], self.s2c_attention[li][0]
            parts = []
            p = c2s.multihead_attn
            W, b = p.in_proj_weight, p.in_proj_bias
            parts += [W[:d].t(), b[:d], W[d:2 * d], W[2 * d:].t(), b[2 * d:], p.out_proj.weight.t(), p.out_proj.bias,
                      c2s.norm.weight, c2s.norm.bias]
            p = c2c.self_attn
            W, b = p.in_proj_weight, p.in_proj_bias
            parts += [W[:d].t(), b[:d], W[d:2 * d].t(), b[d:2 * d], W[2 * d:].t(), b[2 * d:], p.out_proj.weight.t(),
                      p.out_proj.bias, c2c.norm.weight, c2c.norm.bias]
            parts += [ffn.linear1.weight.t(), ffn.linear1.bias, ffn.linear2.weight.t(), ffn.linear2.bias, ffn.norm.weight,
                      ffn.norm.bias]
            p = s2c.multihead_attn
            W, b = p.in_proj_weight, p.in_proj_bias
            parts += [W[d:2 * d].t(), b[d:2 * d], W[2 * d:].t(), b[2 * d:], W[:d], b[:d], p.out_proj.weight.t()]
            m = self.mask_embed_head
            parts += [self.decoder_norm.weight, self.decoder_norm.bias, m[0].weight.t(), m[0].bias, m[2].weight.t(), m[2].bias]
            blob = torch.cat([t.detach().float().contiguous().reshape(-1) for t in parts])
            if blob.numel() != ops.query_blob_floats():
                raise RuntimeError("query weight blob does not match csrc/query_ops.cu")
            cache[1][li] = blob
        return cache[1][li]

    def _dec_params(self):
        ps = getattr(self, "_dec_param_list", None)
        if ps is None:
            ps = self._dec_param_list = [p for n, p in self.named_parameters() if not n.startswith("backbone.")]
        return ps

    @staticmethod
    def _staged_ints(rows, dev):
        """python ints -> device int32 tensor through pinned memory of torch's caching host allocator (an async copy: no
        pageable-copy synchronisation; the allocator keeps the block alive until the copy has run)"""
        t = torch.tensor(rows, dtype=torch.int32)
        if dev.type != "cuda":
            return t.to(dev)
        return t.pin_memory().to(dev, non_blocking=True)

    # ---- the scenes of a batch are independent inside a decoder layer: their voxel-streaming kernels alternate between a few
    #      side streams (forked from / joined into the caller's stream around every phase), so that one scene's kernel
    #      prologue, tail and partial-result merge run under the next scene's kernel instead of between them
    def _side_streams(self, dev, n=None):
        """side streams owned by the CALLER's stream: memory allocated on them is only ever consumed by work that the caller's
        stream orders (fork / join), so two callers on different streams (two batches in flight) never share a pool"""
        n = self.decoder_streams if n is None else n
        if not n:
            return None
        cache = self.__dict__.setdefault("_streams", {})
        key = (dev, ops._raw_stream(), n)
        if key not in cache:
            cache[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
        return cache[key]

    @staticmethod
    def _fork(side):
        if side:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for st in side:
                st.wait_event(ev)

    @staticmethod
    def _join(side):
        if side:
            cur = torch.cuda.current_stream()
            for st in side:
                cur.wait_stream(st)

    @staticmethod
    def _on(side, i):
        import contextlib
        return torch.cuda.stream(side[i % len(side)]) if side else contextlib.nullcontext()

    def _forward_mask_fused(self, pcd_features, coordinates, pos_encodings_pcd, click_idx, click_time_idx):
        H = self.num_heads
        dev = pcd_features.device
        offsets = pcd_features.offsets
        n_scenes = len(offsets) - 1
        tt = self.time_encode.to(dev) if self.time_encode.device != dev else self.time_encode
        self.time_encode = tt
        nbg = self.num_bg_queries
        meta, groups = [], {}
        for b in range(n_scenes):
            ck, ct = click_idx[b], click_time_idx[b]
            K = len(ck) - 1
            split = [len(ck[str(i)]) for i in range(1, K + 1)]
            if min(split, default=1) < 1:
                raise ValueError("every foreground object needs at least one click (agile3d.py:214)")
            n_fg, n_bgc = sum(split), len(ck["0"])
            # query order of a scene: [fg clicks by object id then click order | learned bg | bg clicks] (agile3d.py:249-264)
            src = [offsets[b] + i for o in range(1, K + 1) for i in ck[str(o)]] + [-(k + 1) for k in range(nbg)] \
                + [offsets[b] + i for i in ck["0"]]
            tix = [t for o in range(1, K + 1) for t in ct[str(o)]] + [0] * nbg + list(ct["0"])
            q_obj = [o for o, n in enumerate(split, start=1) for _ in range(n)] + [0] * (nbg + n_bgc)
            meta.append((K, src, tix, q_obj))
            groups.setdefault(n_fg + nbg + n_bgc, []).append(b)
        results = [None] * n_scenes
        pos_list = pos_encodings_pcd[self.hlevels[0]][0].internal
        pos_list_s = pos_encodings_pcd[self.hlevels[0]][0].split
        for nq, members in groups.items():
            B = len(members)
            if nq > ops.S2C_MAX_QUERIES:
                raise ValueError(f"at most {ops.S2C_MAX_QUERIES} click queries per scene, got {nq}")
            ints = self._staged_ints([v for b in members for v in meta[b][1]] + [v for b in members for v in meta[b][2]]
                                     + [b for b in members for _ in range(nq)] + [v for b in members for v in meta[b][3]], dev)
            src_row, time_idx, scene_of_row, q_obj = ints[:B * nq], ints[B * nq:2 * B * nq], ints[2 * B * nq:3 * B * nq], \
                ints[3 * B * nq:].view(B, nq)
            # clicked voxels: xyz is read in the caller's order, the features in the backbone's internal order
            feat_row = src_row if pcd_features.inv is None else \
                torch.where(src_row >= 0, pcd_features.inv[src_row.clamp(min=0).long()], src_row)
            # up to 24 queries per scene the voxel side runs on split rows through the TMA-fed kernels; beyond that on fp32
            # rows (query groups, csrc/decoder_mq.cu), decoded once per scene batch
            split = pcd_features.Fs is not None and pos_list_s is not None and nq <= ops.S2C_SPLIT_MAX_QUERIES
            feats = pcd_features.Fs if split else pcd_features.Fp
            queries, qpos = ops.query_init(feats, coordinates.F, coordinates.range, src_row, time_idx, scene_of_row,
                                           self.pos_enc.gauss_B, tt, self.bg_query_feat.weight, self.bg_query_pos.weight,
                                           feat_row=feat_row if pcd_features.inv is not None else None, feats_split=split)
            srcs = [feats[offsets[b]:offsets[b + 1]] for b in members]
            poss = pos_list_s if split else pos_list
            labels, counts = [None] * B, [None] * B
            outs = [[] for _ in members]
            ctx = torch.empty((B, H * nq, self.hidden_dim), dtype=torch.float32, device=dev)
            side = self._side_streams(dev) if (B > 1 and dev.type == "cuda") else None
            for layer in range(self.num_decoders):
                li = 0 if self.shared_decoder else layer
                s2c = self.s2c_attention[li][0]
                blob = self._layer_blob(li)
                qfold = ops.query_fold_c2s(queries, qpos, blob, B, nq, H)
                last = layer == self.num_decoders - 1
                self._fork(side)
                for i, b in enumerate(members):
                    with self._on(side, i):
                        ops.c2s_attn_fwd(srcs[i], poss[b], qfold[i], nq, H, labels[i], q_obj[i], counts[i], out=ctx[i],
                                         split=split)
                self._join(side)
                q1, qh, kh, vh = ops.query_update_a(ctx, queries, qpos, blob, B, nq, s2c.norm.eps)
                queries, A, c, U, E = ops.query_update_b(q1, qh, kh, vh, qpos, blob, B, nq, H, s2c.norm.eps)
                self._fork(side)
                for i, b in enumerate(members):
                    with self._on(side, i):
                        srcs[i], logits, labels[i], counts[i] = ops.s2c_mask_fwd(
                            srcs[i], poss[b], A[i], c[i], U[i], s2c.multihead_attn.out_proj.bias, s2c.norm.weight,
                            s2c.norm.bias, s2c.norm.eps, E[i], q_obj[i], nq, H, meta[b][0] + 1,
                            x_out=None if layer == 0 else srcs[i],      # never overwrite the caller's backbone features
                            split=split, write_x=not (split and last))  # the last layer's features are never read again
                    outs[i].append(logits)
                self._join(side)
            for i, b in enumerate(members):
                results[b] = [pcd_features.to_caller(b, lg) for lg in outs[i]]      # logits back in the caller's row order
        per_layer = [list(p) for p in zip(*results)]
        out = {"pred_masks": per_layer[-1], "backbone_features": pcd_features}
        if self.aux:
            out["aux_outputs"] = [{"pred_masks": p} for p in per_layer[:-1]]
        return out

    def _forward_mask(self, pcd_features, coordinates, pos_encodings_pcd, click_idx, click_time_idx, grad):
        H, d = self.num_heads, self.hidden_dim
        dev = pcd_features.Fp.device
        offsets = pcd_features.offsets
        n_scenes = len(offsets) - 1
        _WGRAD_TC[0] = self.backbone.algo != ops.ALGO_SIMT
        tt = self.time_encode.to(dev) if self.time_encode.device != dev else self.time_encode
        self.time_encode = tt
        # ---- host side: flatten the click dictionaries (query order per scene: [fg clicks by object id then click
        #      order | 10 learned bg | bg clicks], agile3d.py:249-264) and group scenes by their query count
        meta, groups = [], {}
        for b in range(n_scenes):
            ck, ct = click_idx[b], click_time_idx[b]
            K = len(ck) - 1
            split = [len(ck[str(i)]) for i in range(1, K + 1)]
            if min(split, default=1) < 1:
                raise ValueError("every foreground object needs at least one click (agile3d.py:214)")
            rows = [i for o in range(1, K + 1) for i in ck[str(o)]] + list(ck["0"])
            times = [t for o in range(1, K + 1) for t in ct[str(o)]] + list(ct["0"])
            n_fg, n_bgc = sum(split), len(ck["0"])
            q_obj = [o for o, n in enumerate(split, start=1) for _ in range(n)] + [0] * (self.num_bg_queries + n_bgc)
            meta.append((K, n_fg, n_bgc, rows, times, q_obj))
            groups.setdefault((n_fg, n_bgc), []).append(b)
        results = [None] * n_scenes
        pos_list = pos_encodings_pcd[self.hlevels[0]][0].internal
        for (n_fg, n_bgc), members in groups.items():
            B, nq = len(members), n_fg + self.num_bg_queries + n_bgc
            n_click = n_fg + n_bgc
            grow = torch.tensor([offsets[b] + r for b in members for r in meta[b][3]], dtype=torch.long, device=dev)
            tix = torch.tensor([t for b in members for t in meta[b][4]], dtype=torch.long, device=dev)
            q_obj = torch.tensor([meta[b][5] for b in members], dtype=torch.int32, device=dev)          # [B, nq]
            rng = coordinates.range[torch.tensor(members, device=dev)].repeat_interleave(n_click, dim=0)  # [B*n_click, 6]
            click_feat = pcd_features.Fp[pcd_features.rows_internal(grow)].view(B, n_click, d)
            click_pos = (self._click_pos(coordinates.F[grow], rng[:, :3], rng[:, 3:]) + tt[tix]).view(B, n_click, d)
            bgq = self.bg_query_feat.weight.unsqueeze(0).expand(B, -1, -1)
            bgp = self.bg_query_pos.weight.unsqueeze(0).expand(B, -1, -1)
            queries = torch.cat([click_feat[:, :n_fg], bgq, click_feat[:, n_fg:]], 1)                  # [B, nq, d]
            qpos = torch.cat([click_pos[:, :n_fg], bgp, click_pos[:, n_fg:]], 1)
            srcs = [pcd_features.Fp[offsets[b]:offsets[b + 1]] for b in members]
            labels, counts = [None] * B, [None] * B
            outs = [[] for _ in members]
            ctx = torch.empty((B, H * nq, d), dtype=torch.float32, device=dev)
            for layer in range(self.num_decoders):
                li = 0 if self.shared_decoder else layer
                c2s, c2c = self.c2s_attention[li][0], self.c2c_attention[li][0]
                ffn, s2c = self.ffn_attention[li][0], self.s2c_attention[li][0]
                qfold = self._fold_c2s(c2s.multihead_attn, queries, qpos, H)
                if grad:
                    ctx = torch.stack([_C2sFn.apply(srcs[i], pos_list[b], qfold[i], nq, H, labels[i], q_obj[i], counts[i])
                                       for i, b in enumerate(members)])
                else:
                    for i, b in enumerate(members):
                        ops.c2s_attn_fwd(srcs[i], pos_list[b], qfold[i], nq, H, labels[i], q_obj[i], counts[i],
                                         out=ctx[i])
                q = self._finish_c2s(c2s, queries, ctx, H)
                q = self._self_attn(c2c, q, qpos, H)
                queries = self._ffn(ffn, q)
                A, c, U = self._fold_s2c(s2c.multihead_attn, queries, qpos, H)
                E = self.mask_embed_head(self.decoder_norm(queries)).contiguous()
                for i, b in enumerate(members):
                    if grad:
                        p = s2c.multihead_attn
                        srcs[i], logits, labels[i], counts[i] = _S2cFn.apply(
                            srcs[i], pos_list[b], A[i], c[i], U[i], p.out_proj.bias, s2c.norm.weight, s2c.norm.bias,
                            E[i], s2c.norm.eps, q_obj[i], nq, H, meta[b][0] + 1)
                        outs[i].append(logits)
                        continue
                    srcs[i], logits, labels[i], counts[i] = ops.s2c_mask_fwd(
                        srcs[i], pos_list[b], A[i], c[i], U[i], s2c.multihead_attn.out_proj.bias, s2c.norm.weight,
                        s2c.norm.bias, s2c.norm.eps, E[i], q_obj[i], nq, H, meta[b][0] + 1,
                        x_out=None if layer == 0 else srcs[i])      # never overwrite the caller's backbone features
                    outs[i].append(logits)
            for i, b in enumerate(members):
                results[b] = [pcd_features.to_caller(b, lg) for lg in outs[i]]      # logits back in the caller's row order
        per_layer = [list(p) for p in zip(*results)]
        out = {"pred_masks": per_layer[-1], "backbone_features": pcd_features}
        if self.aux:
            out["aux_outputs"] = [{"pred_masks": p} for p in per_layer[:-1]]
        return out



def build_agile3d(args):
    """models/agile3d.py:399-421 (+ models/backbone.py:5-7)."""
    backbone = Res16UNet34C(3, bn_momentum=args.bn_momentum, conv1_kernel_size=args.conv1_kernel_size)
    return Agile3d(backbone=backbone, hidden_dim=args.hidden_dim, num_heads=args.num_heads,
                   dim_feedforward=args.dim_feedforward, shared_decoder=args.shared_decoder,
                   num_decoders=args.num_decoders, num_bg_queries=args.num_bg_queries, dropout=args.dropout,
                   pre_norm=args.pre_norm, positional_encoding_type=args.positional_encoding_type,
                   normalize_pos_enc=args.normalize_pos_enc, hlevels=args.hlevels, voxel_size=args.voxel_size,
                   gauss_scale=args.gauss_scale, aux=args.aux)
