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


def stem_conv_fwd(coords, feats, table, cap, ksize, weight, out, scale=None, shift=None, relu=True, out_split=False):
    nbr = kernel_map(coords, table, cap, ksize, 1)
    return spconv_fwd(feats, nbr, weight, out, scale, shift, None, relu)


def fourier_posenc(xyz, scene_offsets, gauss_B):
    outs, rng = [], []
    for b in range(len(scene_offsets) - 1):
        p = xyz[scene_offsets[b]:scene_offsets[b + 1]]
        lo, hi = p.min(0)[0], p.max(0)[0]
        t = (((p - lo) / (hi - lo)) * (2 * math.pi)) @ gauss_B
        outs.append(torch.cat([t.sin(), t.cos()], 1))
        rng.append(torch.cat([lo, hi]))
    return torch.cat(outs, 0), torch.stack(rng, 0)


def c2s_attn_fwd(x, pos, qfold, nq, heads, label=None, q_obj=None, obj_count=None, algo=0, out=None):
    s = qfold @ (x + pos).T                                   # [H*nq, Nv]
    if label is not None:
        for r in range(heads * nq):
            o = int(q_obj[r % nq])
            if int(obj_count[o]) > 0:
                s[r, label != o] = float("-inf")
    r = torch.softmax(s, dim=1) @ x
    if out is not None:
        out.copy_(r)
        return out
    return r


def s2c_mask_fwd(x, pos, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out=None, algo=0):
    s = (x + pos) @ A.T + c                                   # [Nv, H*nq]
    a = torch.softmax(s.view(-1, heads, nq), dim=2).reshape(-1, heads * nq)
    y = torch.nn.functional.layer_norm(x + (a @ U + bo), (x.shape[1],), ln_w, ln_b, ln_eps)
    z = y @ E.T                                               # [Nv, nq]
    logits = torch.full((x.shape[0], n_obj), float("-inf"))
    for q in range(nq):
        o = int(q_obj[q])
        logits[:, o] = torch.maximum(logits[:, o], z[:, q])
    label = logits.argmax(1).to(torch.uint8)
    obj_count = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    if x_out is not None:
        x_out.copy_(y)
        y = x_out
    return y, logits, label, obj_count


ALL = ["prepare_tc_weight", "hash_build", "downsample", "kernel_map", "kernel_map_transposed", "spconv_fwd", "stem_conv_fwd",
       "fourier_posenc", "c2s_attn_fwd", "s2c_mask_fwd"]


def patch_ops(monkeypatch):
    import agile3d_b200.ops as ops
    g = globals()
    for name in ALL:
        monkeypatch.setattr(ops, name, g[name])
