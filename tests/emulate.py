"""Torch/numpy emulation of every C-ABI operation's CONTRACT (include/agile3d_b200.h).  TEST-ONLY.

Used by tests/test_host_model_cpu.py to exercise the host-side logic of agile3d_b200 (graph wiring, BatchNorm
folding, concat slices, query folding, click ordering, mask rule) on a CPU-only machine by monkeypatching
``agile3d_b200.ops``.  The product never imports this file.
"""
import math

import numpy as np
import torch

from oracle import me_ref as ME


def _cm(coords):
    cm = ME.CoordinateManager()
    key = ME.CoordinateMapKey(1)
    cm.coords[key] = coords.cpu().numpy().astype(np.int32)
    return cm, key


def hash_build(coords):
    return coords.clone(), 0, torch.zeros(2, dtype=torch.int32)


def downsample(coords, new_stride):
    c = coords.cpu().numpy().astype(np.int64)
    coarse = c.copy()
    coarse[:, 1:] = np.floor_divide(c[:, 1:], new_stride) * new_stride
    k = ME.pack_keys(coarse)
    _, first, inv = np.unique(k, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    out = torch.from_numpy(coarse[np.sort(first)].astype(np.int32))
    parent = torch.from_numpy(rank[inv.reshape(-1)].astype(np.int32))
    return out, out.clone(), 0, parent


def build_levels(coords, n_levels=4, want_offsets=False):
    levels, tables, parents = [coords], [coords.clone()], []
    for lvl in range(n_levels):
        c, t, _, par = downsample(levels[-1], 2 << lvl)
        levels.append(c)
        tables.append(t)
        parents.append(par)
    offsets = None
    if want_offsets:
        b = coords[:, 0].long()
        offsets = [0] + torch.cumsum(torch.bincount(b), 0).tolist()
    return levels, tables, [0] * (1 + n_levels), parents, (0, 0), offsets


def gather_rows(src, idx):
    return src[idx.long()].contiguous()


def brick_rows(coords, parent01, parent12, n_bricks):
    c = coords.long()
    b = parent12.long()[parent01.long()]
    bit = (c[:, 1] & 3) | ((c[:, 2] & 3) << 2) | ((c[:, 3] & 3) << 4)
    rows = torch.full((n_bricks, 64), -1, dtype=torch.int32)
    rows[b, bit] = torch.arange(c.shape[0], dtype=torch.int32)
    return rows


def row_order(nbr, coords):
    K, n = nbr.shape
    mask = ((nbr >= 0).long() << torch.arange(K).unsqueeze(1)).sum(0)
    key = (coords[:, 0].long() << 32) | mask
    perm = torch.sort(key, stable=True)[1]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    return perm.to(torch.int32), inv.to(torch.int32)


def permute_map(nbr, perm_out=None, inv_in=None):
    out = nbr if perm_out is None else nbr[:, perm_out.long()]
    if inv_in is not None:
        out = torch.where(out >= 0, inv_in[out.clamp(min=0).long()], out)
    return out.contiguous()


def kernel_map(out_coords, in_table, cap, ksize, in_tensor_stride, dilation=1, count_pairs=False):
    cm, key = _cm(in_table)
    offs = ME.kernel_offsets(ksize, in_tensor_stride, dilation)
    oc = out_coords.cpu().numpy().astype(np.int64)
    nbr = np.full((offs.shape[0], oc.shape[0]), -1, np.int32)
    for k in range(offs.shape[0]):
        q = oc.copy()
        q[:, 1:] += offs[k]
        nbr[k] = cm.find_rows(key, q)
    nbr = torch.from_numpy(nbr)
    if count_pairs:
        return nbr, (nbr >= 0).sum(1).to(torch.int32)
    return nbr


def kernel_map_transposed(fine_coords, parent, fine_stride):
    f = fine_coords.cpu().numpy().astype(np.int64)
    cs = 2 * fine_stride
    d = (f[:, 1:] - np.floor_divide(f[:, 1:], cs) * cs) // fine_stride
    kidx = d[:, 0] + 2 * d[:, 1] + 4 * d[:, 2]
    nbr = np.full((8, f.shape[0]), -1, np.int32)
    nbr[kidx, np.arange(f.shape[0])] = parent.cpu().numpy()
    return torch.from_numpy(nbr)


def prepare_tc_weight(weight):
    return None


def spconv_fwd(x, nbr, weight, out, scale=None, shift=None, residual=None, relu=False, algo=0, weight_tc=None, **fmt):
    w = weight.detach()
    w = w if w.dim() == 3 else w.unsqueeze(0)
    acc = torch.zeros((out.shape[0], w.shape[2]), dtype=x.dtype)
    for k in range(w.shape[0]):
        if nbr is None:
            acc += x @ w[k]
        else:
            sel = torch.nonzero(nbr[k] >= 0).squeeze(1)
            acc.index_add_(0, sel, x[nbr[k][sel].long()] @ w[k])
    if scale is not None:
        acc = acc * scale
    if shift is not None:
        acc = acc + shift
    if residual is not None:
        acc = acc + residual
    if relu:
        acc = torch.relu(acc)
    out.copy_(acc)
    return out


def stem_conv_fwd(coords, feats, table, cap, ksize, weight, out, scale=None, shift=None, relu=True, out_split=False,
                  bricks=None):
    nbr = kernel_map(coords, table, cap, ksize, 1)
    return spconv_fwd(feats, nbr, weight, out, scale, shift, None, relu)


def unpack_split(xs):
    return xs          # the emulation keeps "split" rows as plain fp32 values (the bf16 hi/lo format is a storage detail)


def fourier_posenc(xyz, scene_offsets, gauss_B, want_split=False):
    if want_split:
        pos, r = fourier_posenc(xyz, scene_offsets, gauss_B)
        return pos, r, pos.clone()
    outs, rng = [], []
    for b in range(len(scene_offsets) - 1):
        p = xyz[scene_offsets[b]:scene_offsets[b + 1]]
        lo, hi = p.min(0)[0], p.max(0)[0]
        t = (((p - lo) / (hi - lo)) * (2 * math.pi)) @ gauss_B
        outs.append(torch.cat([t.sin(), t.cos()], 1))
        rng.append(torch.cat([lo, hi]))
    return torch.cat(outs, 0), torch.stack(rng, 0)


def c2s_attn_fwd(x, pos, qfold, nq, heads, label=None, q_obj=None, obj_count=None, algo=0, out=None, lse=None, split=False):
    s = qfold @ (x + pos).T                                   # [H*nq, Nv]
    if label is not None:
        s = s.clone()
        for r in range(heads * nq):
            o = int(q_obj[r % nq])
            if int(obj_count[o]) > 0:
                s[r, label != o] = float("-inf")
    if lse is not None:
        l = torch.logsumexp(s.detach(), dim=1)
        lse.copy_(torch.where(torch.isinf(l), torch.full_like(l, float("inf")), l))
    r = torch.softmax(s, dim=1) @ x
    if out is not None:
        out.copy_(r)
        return out
    return r


def s2c_mask_fwd(x, pos, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out=None, algo=0, split=False,
                 write_x=True):
    s = (x + pos) @ A.T + c                                   # [Nv, H*nq]
    a = torch.softmax(s.view(-1, heads, nq), dim=2).reshape(-1, heads * nq)
    y = torch.nn.functional.layer_norm(x + (a @ U + bo), (x.shape[1],), ln_w, ln_b, ln_eps)
    z = y @ E.T                                               # [Nv, nq]
    cols = []
    for o in range(n_obj):                                    # autograd-friendly: max over each object's queries
        qs = [q for q in range(nq) if int(q_obj[q]) == o]
        cols.append(z[:, qs].max(dim=1)[0] if qs else torch.full((x.shape[0],), float("-inf"), dtype=z.dtype))
    logits = torch.stack(cols, 1)
    label = logits.argmax(1).to(torch.uint8)
    obj_count = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    if x_out is not None:
        x_out.copy_(y)
        y = x_out
    return (y if write_x else None), logits, label, obj_count


# ------------------------------------------------------------------------------------------------ training step
def bn_stats(z, eps, momentum=0.0, running_mean=None, running_var=None):
    n = z.shape[0]
    mean = z.mean(0)
    var = z.var(0, unbiased=False)
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * var * (n / (n - 1) if n > 1 else 1.0))
    return mean, torch.rsqrt(var + eps)


def bn_apply(z, mean, invstd, gamma, beta, out, residual=None, relu=False):
    y = (z - mean) * invstd * gamma + beta
    if residual is not None:
        y = y + residual
    out.copy_(torch.relu(y) if relu else y)
    return out


def bn_bwd(z, y, dy, mean, invstd, gamma, dz, relu=False, g_out=None):
    n = z.shape[0]
    g = dy * (y > 0) if relu else dy.clone()
    xhat = (z - mean) * invstd
    dbeta, dgamma = g.sum(0), (g * xhat).sum(0)
    res = gamma * invstd * (g - dbeta / n - xhat * dgamma / n)
    if g_out is not None:
        g_out.copy_(g)
    dz.copy_(res)
    return dgamma, dbeta


def col_sum(z):
    return z.sum(0)


def spconv_bwd_weight(x, nbr, dout, K, dweight=None, accumulate=False):
    dw = torch.zeros((K, x.shape[1], dout.shape[1]), dtype=x.dtype)
    for k in range(K):
        if nbr is None:
            dw[k] = x.T @ dout
        else:
            sel = torch.nonzero(nbr[k] >= 0).squeeze(1)
            dw[k] = x[nbr[k][sel].long()].T @ dout[sel]
    if dweight is None:
        return dw
    if accumulate:
        dweight += dw.view_as(dweight)
    else:
        dweight.copy_(dw.view_as(dweight))
    return dweight


def stem_bwd_weight(coords, feats, table, cap, ksize, dz, bricks=None):
    nbr = kernel_map(coords, table, cap, ksize, 1)
    return spconv_bwd_weight(feats, nbr, dz, ksize ** 3)


def decoder_bwd_rows(nq, heads):
    for J in (6, 8, 10, 12, 14, 16):
        if nq * heads <= 16 * J:
            return 16 * J
    raise RuntimeError("too many queries")


def c2s_attn_bwd(x, pos, qf, qft, dctx, dctxt, lse, dr, rowobj, hqp, label):
    assert torch.equal(qft, qf.T) and torch.equal(dctxt, dctx.T) and qf.shape[0] == hqp
    S = (x + pos) @ qf.T                                      # [Nv, hqp]
    P = torch.exp(S - lse)
    dead = (rowobj == -2).unsqueeze(0).expand_as(P).clone()
    if label is not None:
        dead |= (rowobj >= 0).unsqueeze(0) & (label.long().unsqueeze(1) != rowobj.long().unsqueeze(0))
    P = torch.where(dead, torch.zeros_like(P), P)
    dS = P * (x @ dctx.T - dr)
    return P @ dctx + dS @ qf, dS


def s2c_mask_bwd(x, pos, A, At, c, U, Ut, bo, ln_w, ln_b, ln_eps, E, Et, q_obj, nq, heads, n_obj, hqp, dxo, dlogits):
    assert torch.equal(At, A.T) and torch.equal(Ut, U.T) and torch.equal(Et, E.T) and A.shape[0] == hqp
    nv, HQ = x.shape[0], heads * nq
    xp = x + pos
    a = torch.zeros((nv, hqp), dtype=x.dtype)
    a[:, :HQ] = torch.softmax((xp @ A[:HQ].T + c[:HQ]).view(nv, heads, nq), dim=2).reshape(nv, HQ)
    y = x + a @ U + bo
    mu, var = y.mean(1, keepdim=True), y.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + ln_eps)
    n = (y - mu) * rstd
    xo = n * ln_w + ln_b
    prods = xo @ E[:nq].T
    g = torch.zeros((nv, 32), dtype=x.dtype)
    if dlogits is not None:
        for o in range(n_obj):
            cols = [q for q in range(nq) if int(q_obj[q]) == o]
            if cols:
                first = prods[:, cols].argmax(1)              # torch.argmax returns the first maximal index
                g[torch.arange(nv), torch.tensor(cols)[first]] = dlogits[:, o]
    t = g @ E
    if dxo is not None:
        t = t + dxo
    dn = t * ln_w
    dy = rstd * (dn - dn.mean(1, keepdim=True) - n * (dn * n).mean(1, keepdim=True))
    da = dy @ U.T
    ds = torch.zeros_like(a)
    a3, da3 = a[:, :HQ].view(nv, heads, nq), da[:, :HQ].view(nv, heads, nq)
    ds[:, :HQ] = (a3 * (da3 - (a3 * da3).sum(2, keepdim=True))).reshape(nv, HQ)
    dx = dy + ds @ A
    cols = torch.cat([dy.sum(0), (t * n).sum(0), t.sum(0), ds.sum(0)])
    return dx, a, ds, dy, g, cols


def loss_fwd(logits, target, w, eps=1e-6):
    C = logits.shape[1]
    lse = torch.logsumexp(logits, 1)
    lt = logits.gather(1, target.long().unsqueeze(1)).squeeze(1)
    pt = torch.exp(lt - lse)
    num, den = 2 * pt / C, 2.0 / C
    dice = torch.where(num > eps, 1 - (num + eps) / (den + eps), torch.zeros_like(num))
    return torch.stack([(w * (lse - lt)).sum(), (w * dice).sum()])


def loss_bwd(logits, target, w, g, eps=1e-6):
    n, C = logits.shape
    p = torch.softmax(logits, 1)
    ind = torch.nn.functional.one_hot(target.long(), C).to(p.dtype)
    pt = (p * ind).sum(1, keepdim=True)
    den = 2.0 / C
    kd = torch.where(2 * pt / C > eps, -(den / (den + eps)) * pt * g[1], torch.zeros_like(pt))
    return (w / n).unsqueeze(1) * (g[0] * (p - ind) + kd * (ind - p))


def click_loss_weights(xyz, clicks, alpha=0.8, beta=2.0, tita=0.3):
    d = torch.cdist(xyz, clicks).min(1)[0]
    return alpha + (beta - alpha) * (1 - torch.clamp(d, max=tita) / tita)


def grad_norm(flat):
    return flat.norm().reshape(1)


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, norm=None, max_norm=0.0):
    coef = 1.0
    if norm is not None and max_norm > 0:
        coef = min(1.0, max_norm / (float(norm) + 1e-6))
    gi = g * coef
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_((1 - beta1) * gi)
    v.mul_(beta2).add_((1 - beta2) * gi * gi)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p.sub_((lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps))


def wgrad_tc_supported(K, cin, cout):
    return False          # the emulation has one weight-gradient path (the host logic then never asks for split rows)


def c2s_attn_bwd_tc(x, pos, qf, dctx, lse, dr, rowobj, hqp, label):
    """same contract as c2s_attn_bwd, dS returned in column chunks of at most 256 (ops.c2s_attn_bwd_tc)"""
    dx, ds = c2s_attn_bwd(x, pos, qf, qf.t().contiguous(), dctx, dctx.t().contiguous(), lse, dr, rowobj, hqp, label)
    return dx, [ds[:, a:a + 256].contiguous() for a in range(0, hqp, 256)]


def s2c_mask_bwd_tc_any(x, pos, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, dxo, dlogits, x_out, xt_dy=None):
    """gradients of s2c_mask_fwd by torch.autograd on its emulation (any number of queries)
    -> (dx, dA, dc, dU, dbo, dln_w, dln_b, dE)"""
    leaves = [t.detach().clone().requires_grad_(True) for t in (x, A, c, U, bo, ln_w, ln_b, E)]
    with torch.enable_grad():
        xl, Al, cl, Ul, bol, lwl, lbl, El = leaves
        y, logits, _, _ = s2c_mask_fwd(xl, pos, Al, cl, Ul, bol, lwl, lbl, ln_eps, El, q_obj, nq, heads, n_obj)
        loss = 0
        if dxo is not None:
            loss = loss + (y * dxo).sum()
        if dlogits is not None:
            fin = torch.isfinite(logits)
            loss = loss + (torch.where(fin, logits, torch.zeros_like(logits)) * dlogits).sum()
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    return tuple(g if g is not None else torch.zeros_like(l) for g, l in zip(grads, leaves))


# ------------------------------------------------------------------------------------------------ click-query side (K11)
# Contract emulations of ag3d_query_* (include/agile3d_b200.h).  The weight blob is decoded here from the layout
# documented in csrc/query_ops.cu, independently of agile3d_b200/model.py::_layer_blob that builds it.
_QD, _QF = 128, 1024
_BLOB_FIELDS = [("c2s_WqT", (128, 128)), ("c2s_bq", (128,)), ("c2s_Wk", (128, 128)), ("c2s_WvT", (128, 128)), ("c2s_bv", (128,)),
                ("c2s_WoT", (128, 128)), ("c2s_bo", (128,)), ("c2s_lnw", (128,)), ("c2s_lnb", (128,)),
                ("c2c_WqT", (128, 128)), ("c2c_bq", (128,)), ("c2c_WkT", (128, 128)), ("c2c_bk", (128,)), ("c2c_WvT", (128, 128)),
                ("c2c_bv", (128,)), ("c2c_WoT", (128, 128)), ("c2c_bo", (128,)), ("c2c_lnw", (128,)), ("c2c_lnb", (128,)),
                ("ffn_W1T", (128, 1024)), ("ffn_b1", (1024,)), ("ffn_W2T", (1024, 128)), ("ffn_b2", (128,)), ("ffn_lnw", (128,)),
                ("ffn_lnb", (128,)),
                ("s2c_WkT", (128, 128)), ("s2c_bk", (128,)), ("s2c_WvT", (128, 128)), ("s2c_bv", (128,)), ("s2c_Wq", (128, 128)),
                ("s2c_bq", (128,)), ("s2c_WoT", (128, 128)),
                ("dec_lnw", (128,)), ("dec_lnb", (128,)), ("M1T", (128, 128)), ("m1b", (128,)), ("M2T", (128, 128)), ("m2b", (128,))]


def query_blob_floats():
    return sum(int(np.prod(sh)) for _, sh in _BLOB_FIELDS)


def _blob(blob):
    out, off = {}, 0
    for name, sh in _BLOB_FIELDS:
        n = int(np.prod(sh))
        out[name] = blob[off:off + n].view(*sh)
        off += n
    assert off == blob.numel()
    return out


def _ln(x, w, b, eps):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)


def query_init(feats, xyz, rng, src_row, time_idx, scene_of_row, gauss_B, time_table, bg_feat, bg_pos, feat_row=None,
               feats_split=False):
    n = src_row.shape[0]
    q = torch.empty((n, _QD), dtype=feats.dtype)
    qp = torch.empty_like(q)
    for r in range(n):
        s = int(src_row[r])
        if s < 0:
            q[r], qp[r] = bg_feat[-(s + 1)], bg_pos[-(s + 1)]
        else:
            lo, hi = rng[int(scene_of_row[r]), :3], rng[int(scene_of_row[r]), 3:]
            t = (((xyz[s] - lo) / (hi - lo)) * (2 * math.pi)) @ gauss_B
            q[r] = feats[s if feat_row is None else int(feat_row[r])]
            qp[r] = torch.cat([t.sin(), t.cos()]) + time_table[int(time_idx[r])]
    return q, qp


def query_fold_c2s(queries, qpos, blob, B, nq, heads=8):
    w = _blob(blob)
    qp = (((queries + qpos).view(B, nq, _QD) @ w["c2s_WqT"]) + w["c2s_bq"]) * 0.25
    return torch.einsum("bqhd,hdc->bhqc", qp.view(B, nq, heads, 16), w["c2s_Wk"].view(heads, 16, _QD)).reshape(B, heads * nq, _QD).contiguous()


def query_update_a(ctx, queries, qpos, blob, B, nq, ln_eps=1e-5):
    w = _blob(blob)
    H = 8
    Q, P = queries.view(B, nq, _QD), qpos.view(B, nq, _QD)
    WvT = w["c2s_WvT"].view(_QD, H, 16)                                            # [c, h, d]
    heads = torch.einsum("bhqc,chd->bqhd", ctx.view(B, H, nq, _QD), WvT).reshape(B, nq, _QD) + w["c2s_bv"]
    q1 = _ln(Q + (heads @ w["c2s_WoT"] + w["c2s_bo"]), w["c2s_lnw"], w["c2s_lnb"], ln_eps)
    x = q1 + P
    return q1, x @ w["c2c_WqT"] + w["c2c_bq"], x @ w["c2c_WkT"] + w["c2c_bk"], q1 @ w["c2c_WvT"] + w["c2c_bv"]


def query_update_b(q1, qh, kh, vh, qpos, blob, B, nq, heads=8, ln_eps=1e-5):
    w = _blob(blob)
    H = heads
    P = qpos.view(B, nq, _QD)
    sp = lambda t: t.reshape(B, nq, H, 16).transpose(1, 2)                          # [B, H, nq, 16]
    a = torch.softmax((sp(qh) * 0.25) @ sp(kh).transpose(2, 3), dim=-1)
    o = (a @ sp(vh)).transpose(1, 2).reshape(B, nq, _QD)
    q2 = _ln(q1 + (o @ w["c2c_WoT"] + w["c2c_bo"]), w["c2c_lnw"], w["c2c_lnb"], ln_eps)
    q3 = _ln(q2 + (torch.relu(q2 @ w["ffn_W1T"] + w["ffn_b1"]) @ w["ffn_W2T"] + w["ffn_b2"]), w["ffn_lnw"], w["ffn_lnb"], ln_eps)
    kp = ((q3 + P) @ w["s2c_WkT"] + w["s2c_bk"]).view(B, nq, H, 16)
    vp = (q3 @ w["s2c_WvT"] + w["s2c_bv"]).view(B, nq, H, 16)
    A = (torch.einsum("bqhd,hdc->bhqc", kp, w["s2c_Wq"].view(H, 16, _QD)) * 0.25).reshape(B, H * nq, _QD).contiguous()
    c = (torch.einsum("bqhd,hd->bhq", kp, w["s2c_bq"].view(H, 16)) * 0.25).reshape(B, H * nq).contiguous()
    U = torch.einsum("bqhd,hdc->bhqc", vp, w["s2c_WoT"].view(H, 16, _QD)).reshape(B, H * nq, _QD).contiguous()
    E = torch.relu(_ln(q3, w["dec_lnw"], w["dec_lnb"], ln_eps) @ w["M1T"] + w["m1b"]) @ w["M2T"] + w["m2b"]
    return q3, A, c, U, E


ALL = ["wgrad_tc_supported", "c2s_attn_bwd_tc", "s2c_mask_bwd_tc_any",
       "bn_stats", "bn_apply", "bn_bwd", "col_sum", "spconv_bwd_weight", "stem_bwd_weight", "decoder_bwd_rows",
       "c2s_attn_bwd", "s2c_mask_bwd", "loss_fwd", "loss_bwd", "click_loss_weights", "grad_norm", "adamw_step",
       "prepare_tc_weight", "hash_build", "downsample", "build_levels", "unpack_split", "gather_rows", "brick_rows", "row_order", "permute_map", "kernel_map", "kernel_map_transposed", "spconv_fwd", "stem_conv_fwd",
       "fourier_posenc", "c2s_attn_fwd", "s2c_mask_fwd", "query_blob_floats", "query_init", "query_fold_c2s",
       "query_update_a", "query_update_b"]


def patch_ops(monkeypatch):
    import agile3d_b200.ops as ops
    g = globals()
    for name in ALL:
        monkeypatch.setattr(ops, name, g[name])
