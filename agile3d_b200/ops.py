"""Thin torch-tensor wrappers over the C-ABI (include/agile3d_b200.h).  torch is plumbing here: it owns the
device memory and the stream; every operation below is one call into libagile3d_b200.so."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_TC_PACKED, RELU, check, lib  # noqa: F401

SLOT_BYTES = 16


class Profiler:
    """Per-family CUDA-event timing + algorithmic bytes/flops, for bench.py's roofline line.  Events are recorded on
    torch's current stream, which is the stream every kernel of this library is launched on."""

    def __init__(self):
        self.records = []       # (family, algorithmic_bytes, flops, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        fam = {}
        for name, nbytes, flops, e0, e1 in self.records:
            f = fam.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            f["launches"] += 1
            f["ms"] += e0.elapsed_time(e1)
            f["bytes"] += int(nbytes)
            f["flops"] += int(flops)
        return fam


_prof = None


def set_profiler(p):
    global _prof
    _prof = p


class _Timed:
    def __init__(self, family, nbytes, flops=0):
        self.args = (family, nbytes, flops)

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None and exc[0] is None:
            self.e1.record()
            f, b, fl = self.args
            _prof.records.append((f, b() if callable(b) else b, fl() if callable(fl) else fl, self.e0, self.e1))
        return False


class _NoTimed:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_TIMED = _NoTimed()
_TimedOn = _Timed


def _Timed(family, nbytes, flops=0):      # noqa: N802 - keeps the call sites; no object, no events unless a profiler is set
    return _TimedOn(family, nbytes, flops) if _prof is not None else _NO_TIMED


_ws_cache = {}
_spconv_ws_bytes = {}

# Generation counter of the model parameters: bumped by every writer that changes parameter VALUES without going
# through torch's version counter (FlatAdamW.step's raw kernel, load_state_dict, .to()).  Caches of derived weight
# images (folded BatchNorm, bf16 hi/lo tensor-core images, transposed kernels) key on it.
_param_generation = [0]


def bump_param_generation():
    _param_generation[0] += 1


def param_generation():
    return _param_generation[0]


def _raw_stream():
    """raw cudaStream_t of torch's current stream on the current device (two C calls; torch.cuda.current_stream() builds
    a Stream object and costs several microseconds per library call)"""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _workspace(tag, device, nbytes):
    """Grow-only scratch buffer per (purpose, device, stream); stream-ordered reuse is safe on one stream."""
    key = (tag, device, _raw_stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _ws_cache[key] = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
    return buf


def kernel_launches():
    return int(lib().ag3d_kernel_launches())


def _stream():
    return C.c_void_p(_raw_stream())


def _p(t):
    return t.data_ptr() if t is not None else None          # ctypes converts int / None to void* (argtypes are set)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.Ag3dError("agile3d_b200 ops need CUDA tensors (there is no CPU fallback)")


def _rows2d(t, dtype=torch.float32):
    """2-D tensor whose rows are contiguous (a channel slice of a wider buffer is fine) -> (ptr, ld)."""
    if t.dim() != 2 or t.dtype != dtype or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.Ag3dError(f"expected a row-contiguous 2-D {dtype} tensor, got {tuple(t.shape)} {t.dtype} {t.stride()}")
    return _p(t), int(t.stride(0))


def hash_build(coords):
    """coords int32 [N,4] -> (table uint8[cap*16], cap).  Raises on duplicate / out-of-range coordinates."""
    _need_cuda(coords)
    n = coords.shape[0]
    cap = lib().ag3d_hash_capacity(n)
    table = torch.empty(cap * SLOT_BYTES, dtype=torch.uint8, device=coords.device)
    status = torch.zeros(2, dtype=torch.int32, device=coords.device)
    with _Timed("maps", 16 * n + 16 * cap):
        check(lib().ag3d_hash_build(_p(coords), n, _p(table), cap, _p(status), _stream()), "ag3d_hash_build")
    return table, cap, status


def downsample(coords, new_stride):
    """-> (coarse coords int32 [M,4], coarse table, cap, parent int32 [N]).  One host sync (reads M)."""
    _need_cuda(coords)
    n = coords.shape[0]
    dev = coords.device
    cap = lib().ag3d_hash_capacity(n)
    table = torch.empty(cap * SLOT_BYTES, dtype=torch.uint8, device=dev)
    parent = torch.empty(n, dtype=torch.int32, device=dev)
    out = torch.empty((n, 4), dtype=torch.int32, device=dev)
    out_n = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = lib().ag3d_downsample_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with _Timed("maps", 16 * n + 16 * cap + 4 * n):
        check(lib().ag3d_downsample(_p(coords), n, new_stride, _p(table), cap, _p(parent), _p(out), _p(out_n),
                                    _p(ws), wsb, _stream()), "ag3d_downsample")
    m = int(out_n.item())
    return out[:m], table, cap, parent


MAX_SCENES = 256       # scenes per batched coordinate list (scene_offsets)


def build_levels(coords, n_levels=4, want_offsets=False):
    """Level-0 hash table + `n_levels` chained stride-2 levels (+ the scene row ranges) with ONE host read-back: the row
    count of level l stays on the device and feeds the kernels of level l+1 (ag3d_downsample_dev).
    -> (coords[1+n], tables[1+n], caps[1+n], parents[n], (dup, out_of_range), offsets list | None)"""
    _need_cuda(coords)
    n0, dev = coords.shape[0], coords.device
    cap = lib().ag3d_hash_capacity(n0)
    i32 = dict(dtype=torch.int32, device=dev)
    meta = torch.zeros(2 + n_levels + MAX_SCENES + 2, **i32)        # status[2] | counts[n_levels] | offsets[MAX_SCENES + 2]
    tables = [torch.empty(cap * SLOT_BYTES, dtype=torch.uint8, device=dev) for _ in range(1 + n_levels)]
    with _Timed("maps", 16 * n0 + 16 * cap):
        check(lib().ag3d_hash_build(_p(coords), n0, _p(tables[0]), cap, _p(meta), _stream()), "ag3d_hash_build")
    if want_offsets:
        check(lib().ag3d_scene_offsets(_p(coords), n0, MAX_SCENES, _p(meta[2 + n_levels:]), _stream()), "ag3d_scene_offsets")
    wsb = lib().ag3d_downsample_workspace_bytes(n0)
    ws = _workspace("downsample", dev, wsb)
    bufs, parents = [coords], []
    for lvl in range(n_levels):
        out = torch.empty((n0, 4), **i32)
        parent = torch.empty(n0, **i32)
        cnt = meta[2 + lvl:3 + lvl]
        with _Timed("maps", 16 * n0 + 16 * cap + 4 * n0):
            if lvl == 0:
                check(lib().ag3d_downsample(_p(bufs[lvl]), n0, 2 << lvl, _p(tables[lvl + 1]), cap, _p(parent), _p(out), _p(cnt),
                                            _p(ws), ws.numel(), _stream()), "ag3d_downsample")
            else:
                check(lib().ag3d_downsample_dev(_p(bufs[lvl]), n0, _p(meta[1 + lvl:2 + lvl]), 2 << lvl, _p(tables[lvl + 1]), cap,
                                                _p(parent), _p(out), _p(cnt), _p(ws), ws.numel(), _stream()),
                      "ag3d_downsample_dev")
        bufs.append(out)
        parents.append(parent)
    host = meta.cpu().tolist()                                       # the one synchronisation of the map build
    sizes = [n0] + host[2:2 + n_levels]
    levels = [bufs[l][:sizes[l]] for l in range(1 + n_levels)]
    parents = [parents[l][:sizes[l]] for l in range(n_levels)]
    offsets = None
    if want_offsets:
        offs = host[2 + n_levels:]
        if offs[MAX_SCENES + 1]:
            raise ValueError("rows of a scene must be contiguous and scenes in batch order (batched_coordinates)")
        if offs[MAX_SCENES] != n0:
            raise ValueError(f"more than {MAX_SCENES} scenes in one batched coordinate list")
        n_scenes = max(b for b in range(MAX_SCENES) if offs[b] < n0) + 1
        offsets = offs[:n_scenes] + [n0]
    return levels, tables, [cap] * (1 + n_levels), parents, (host[0], host[1]), offsets


def row_order(nbr, coords):
    """Internal row order of a level: rows sorted scene by scene by their neighbour pattern -> (perm [new] = old row,
    inv [old] = new row), int32 (csrc/coords.cu: ag3d_row_order)."""
    _need_cuda(nbr, coords)
    K, n = nbr.shape
    perm = torch.empty(n, dtype=torch.int32, device=nbr.device)
    inv = torch.empty(n, dtype=torch.int32, device=nbr.device)
    wsb = lib().ag3d_row_order_workspace_bytes(n)
    ws = _workspace("row_order", nbr.device, wsb)
    with _Timed("maps", 4 * K * n + 16 * n + 8 * n):
        check(lib().ag3d_row_order(_p(nbr), K, n, _p(coords), _p(perm), _p(inv), _p(ws), ws.numel(), _stream()), "ag3d_row_order")
    return perm, inv


def permute_map(nbr, perm_out=None, inv_in=None):
    """out[k][i] = inv_in[nbr[k][perm_out[i]]] (-1 stays; None = identity)."""
    _need_cuda(nbr)
    K, n = nbr.shape
    out = torch.empty_like(nbr)
    with _Timed("maps", 8 * K * n):
        check(lib().ag3d_permute_map(_p(nbr), K, n, _p(perm_out), _p(inv_in), _p(out), _stream()), "ag3d_permute_map")
    return out


def kernel_map(out_coords, in_table, cap, ksize, in_tensor_stride, dilation=1, count_pairs=False):
    """-> nbr int32 [K, N_out] (and pair counts int32 [K] if asked)."""
    _need_cuda(out_coords, in_table)
    n_out = out_coords.shape[0]
    K = ksize ** 3
    nbr = torch.empty((K, n_out), dtype=torch.int32, device=out_coords.device)
    pc = torch.zeros(K, dtype=torch.int32, device=out_coords.device) if count_pairs else None
    with _Timed("maps", 16 * n_out + 16 * K * n_out + 4 * K * n_out):     # query read + probes + table write
        check(lib().ag3d_kernel_map(_p(out_coords), n_out, _p(in_table), cap, ksize, in_tensor_stride, dilation,
                                    _p(nbr), _p(pc), _stream()), "ag3d_kernel_map")
    return (nbr, pc) if count_pairs else nbr


def kernel_map_transposed(fine_coords, parent, fine_stride):
    _need_cuda(fine_coords, parent)
    n = fine_coords.shape[0]
    nbr = torch.empty((8, n), dtype=torch.int32, device=fine_coords.device)
    with _Timed("maps", 20 * n + 32 * n):
        check(lib().ag3d_kernel_map_transposed(_p(fine_coords), _p(parent), n, fine_stride, _p(nbr), _stream()),
              "ag3d_kernel_map_transposed")
    return nbr


def prepare_tc_weight(weight):
    """fp32 [K,cin,cout] (or [cin,cout]) -> opaque uint8 tensor holding the bf16 hi/lo stage images of the tcgen05 path."""
    _need_cuda(weight)
    w = weight.detach()
    w = (w if w.dim() == 3 else w.unsqueeze(0)).contiguous().float()
    K, cin, cout = w.shape
    nbytes = lib().ag3d_spconv_tc_weight_bytes(K, cin, cout)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib().ag3d_spconv_tc_prepare_weight(_p(w), K, cin, cout, _p(buf), _stream()), "ag3d_spconv_tc_prepare_weight")
    return buf


IN_SPLIT, OUT_SPLIT, RES_SPLIT = 2, 4, 8


def pack_split(x):
    """fp32 [N,C] -> "split" rows (64 B bf16 hi | 64 B bf16 lo per 32-channel slab) viewed as fp32 [N,C] (same bytes)."""
    n, c = x.shape
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    both = torch.stack([hi.view(n, c // 32, 32), lo.view(n, c // 32, 32)], dim=2)    # [N, C/32, 2, 32] bf16
    return both.contiguous().view(torch.int16).view(n, c * 2).view(torch.float32).view(n, c)


def unpack_split(xs):
    """inverse of pack_split (up to the 2^-17 relative truncation): split rows [N,C] -> fp32 [N,C]."""
    n, c = xs.shape
    both = xs.contiguous().view(torch.int16).view(n, c // 32, 2, 32).view(torch.bfloat16).float()
    return (both[:, :, 0] + both[:, :, 1]).reshape(n, c)


def spconv_fwd(x, nbr, weight, out, scale=None, shift=None, residual=None, relu=False, algo=ALGO_AUTO, weight_tc=None,
               in_split=False, out_split=False, res_split=False):
    """out[o] = act(scale * sum_k x[nbr[k][o]] @ weight[k] + shift (+ residual[o])).  x/out/residual may be channel
    slices of wider buffers.  weight [K,cin,cout] (or [cin,cout] with nbr None)."""
    _need_cuda(x, weight, out)
    xp, x_ld = _rows2d(x)
    op, o_ld = _rows2d(out)
    rp, r_ld = (_rows2d(residual) if residual is not None else (C.c_void_p(0), 0))
    w = weight if weight.dim() == 3 else weight.unsqueeze(0)
    K, cin, cout = w.shape
    if not w.is_contiguous() or w.dtype != torch.float32:
        raise _lib.Ag3dError("weight must be contiguous fp32 [K,cin,cout]")
    if x.shape[1] != cin or out.shape[1] != cout:
        raise _lib.Ag3dError(f"channel mismatch: x {tuple(x.shape)}, w {tuple(w.shape)}, out {tuple(out.shape)}")
    n_out = out.shape[0]
    if nbr is not None and (tuple(nbr.shape) != (K, n_out) or not nbr.is_contiguous()):
        raise _lib.Ag3dError(f"neighbour table must be contiguous int32 [{K},{n_out}], got {tuple(nbr.shape)}")
    if nbr is None and x.shape[0] != n_out:
        raise _lib.Ag3dError("1x1 conv needs as many input as output rows")
    pairs = n_out
    if _prof is not None and nbr is not None:
        pairs = int((nbr >= 0).sum().item())
    # 4*N_in*Cin + 4*N_out*Cout + 8*P + 4*K*Cin*Cout (+ 4*N_out*Cout residual); flops 2*P*Cin*Cout
    nbytes = 4 * x.shape[0] * cin + 4 * n_out * cout + 8 * pairs + 4 * K * cin * cout \
        + (4 * n_out * cout if residual is not None else 0)
    flags = (RELU if relu else 0) | (IN_SPLIT if in_split else 0) | (OUT_SPLIT if out_split else 0) \
        | (RES_SPLIT if (res_split and residual is not None) else 0)
    ws, wsb = None, 0
    if weight_tc is not None and algo != ALGO_SIMT:
        wkey = (n_out, K, cin, cout)
        wsb = _spconv_ws_bytes.get(wkey)
        if wsb is None:
            wsb = _spconv_ws_bytes[wkey] = lib().ag3d_spconv_workspace_bytes(n_out, K, cin, cout)
            if len(_spconv_ws_bytes) > 4096:
                _spconv_ws_bytes.clear()
        if wsb:
            ws = _workspace("spconv", x.device, wsb)
    with _Timed("spconv", nbytes, 2 * pairs * cin * cout):
        rc = lib().ag3d_spconv_fwd_rows(xp, x.shape[0], x_ld, cin, _p(nbr), K, n_out, _p(w), _p(weight_tc), cout,
                                        _p(scale), _p(shift), rp, r_ld, op, o_ld, flags, algo, _p(ws), wsb, _stream())
        if rc:
            check(rc, f"ag3d_spconv_fwd_rows(x {tuple(x.shape)} ld {x_ld} @{x.data_ptr():#x}, K {K}, out {tuple(out.shape)} ld {o_ld} "
                      f"@{out.data_ptr():#x}, residual {None if residual is None else hex(residual.data_ptr())}, flags {flags}, "
                      f"algo {algo}, ws {wsb})")
    return out


def gather_rows(src, idx):
    """src[idx] for a contiguous 2-D tensor with 4-byte elements and an int32 index vector (row permutations)"""
    _need_cuda(src, idx)
    if src.dim() != 2 or not src.is_contiguous() or src.element_size() != 4 or idx.dtype != torch.int32:
        raise _lib.Ag3dError("gather_rows: contiguous 2-D tensor of 4-byte elements and an int32 index")
    out = torch.empty((idx.shape[0], src.shape[1]), dtype=src.dtype, device=src.device)
    if idx.shape[0]:
        check(lib().ag3d_gather_rows(_p(src), 4 * src.shape[1], _p(idx), idx.shape[0], _p(out), _stream()), "ag3d_gather_rows")
    return out


def brick_rows(coords, parent01, parent12, n_bricks):
    """full-resolution row of every cell of every tensor-stride-4 voxel: int32 [n_bricks, 64], -1 = empty (csrc/common.cuh)"""
    _need_cuda(coords, parent01, parent12)
    n = coords.shape[0]
    rows = torch.empty((n_bricks, 64), dtype=torch.int32, device=coords.device)
    with _Timed("maps", 16 * n + 8 * n + 256 * n_bricks):
        check(lib().ag3d_brick_rows(_p(coords), _p(parent01), _p(parent12), n, n_bricks, _p(rows), _stream()),
              "ag3d_brick_rows")
    return rows


def stem_conv_fwd(coords, feats, table, cap, ksize, weight, out, scale=None, shift=None, relu=True, out_split=False,
                  bricks=None):
    """bricks = (tensor-stride-4 table, its capacity, brick_rows): the neighbours come from 8 probes + the bricks' row
    lists instead of ksize^3 probes of the full-resolution table (same result)."""
    _need_cuda(coords, feats, table, weight, out)
    if feats.shape[1] != 3 or weight.shape[-2:] != (3, 32) or out.shape[1] != 32:
        raise _lib.Ag3dError("stem conv is 3 -> 32 channels")
    if not (feats.is_contiguous() and weight.is_contiguous()):
        raise _lib.Ag3dError("stem inputs must be contiguous")
    op, o_ld = _rows2d(out)
    n, K = coords.shape[0], ksize ** 3
    flags = (RELU if relu else 0) | (OUT_SPLIT if out_split else 0)
    with _Timed("stem", 16 * n + 12 * n + 4 * n * 32 + 16 * K * n + 4 * K * 96):
        if bricks is not None and ksize in (3, 5):
            t2, cap2, rows = bricks
            check(lib().ag3d_stem_conv_fwd_bricks(_p(coords), _p(feats), n, _p(t2), cap2, _p(rows), ksize, _p(weight),
                                                  _p(scale), _p(shift), op, o_ld, flags, _stream()),
                  "ag3d_stem_conv_fwd_bricks")
        else:
            check(lib().ag3d_stem_conv_fwd(_p(coords), _p(feats), n, _p(table), cap, ksize, _p(weight),
                                           _p(scale), _p(shift), op, o_ld, flags, _stream()),
                  "ag3d_stem_conv_fwd")
    return out


def fourier_posenc(xyz, scene_offsets, gauss_B, want_split=False):
    """xyz f32 [N,3] (scenes contiguous), scene_offsets python list of B+1 ints -> (pos f32 [N,d], range f32 [B,6]);
    want_split: -> (pos, range, pos as "split" rows) for the TMA-fed decoder kernels."""
    _need_cuda(xyz, gauss_B)
    nb = len(scene_offsets) - 1
    d = 2 * gauss_B.shape[1]
    out = torch.empty((xyz.shape[0], d), dtype=torch.float32, device=xyz.device)
    rng = torch.empty((nb, 6), dtype=torch.float32, device=xyz.device)
    wsb = lib().ag3d_posenc_workspace_bytes(nb)
    ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
    offs = (C.c_int32 * (nb + 1))(*scene_offsets)
    if want_split:
        outs = torch.empty_like(out)
        with _Timed("posenc", 2 * 12 * xyz.shape[0] + 4 * d * xyz.shape[0]):
            check(lib().ag3d_fourier_posenc_split(_p(xyz.contiguous()), C.cast(offs, C.c_void_p), nb, _p(gauss_B.contiguous()),
                                                  d, _p(out), _p(outs), _p(rng), _p(ws), wsb, _stream()),
                  "ag3d_fourier_posenc_split")
        return out, rng, outs
    with _Timed("posenc", 2 * 12 * xyz.shape[0] + 4 * d * xyz.shape[0]):
        check(lib().ag3d_fourier_posenc(_p(xyz.contiguous()), C.cast(offs, C.c_void_p), nb, _p(gauss_B.contiguous()),
                                        d, _p(out), _p(rng), _p(ws), wsb, _stream()), "ag3d_fourier_posenc")
    return out, rng


_c2s_ws = {}


def c2s_attn_fwd(x, pos, qfold, nq, heads, label=None, q_obj=None, obj_count=None, algo=ALGO_AUTO, out=None, lse=None,
                 split=False):
    """-> ctx f32 [heads*nq, 128].  lse (optional f32 [heads*nq]) receives the rows' log-sum-exp (for the backward).
    split=True: x and pos are "split" rows (pack_split_rows / the tensor-core backbone's format), streamed by the TMA engine."""
    _need_cuda(x, pos, qfold)
    for t in (x, pos, qfold):
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise _lib.Ag3dError("c2s inputs must be contiguous fp32")
    ctx = out if out is not None else torch.empty((heads * nq, x.shape[1]), dtype=torch.float32, device=x.device)
    if tuple(ctx.shape) != (heads * nq, x.shape[1]) or not ctx.is_contiguous():
        raise _lib.Ag3dError("c2s output must be contiguous [heads*nq, 128]")
    wsb = lib().ag3d_c2s_workspace_bytes(nq, heads)
    key = (x.device, _raw_stream())
    ws = _c2s_ws.get(key)
    if ws is None or ws.numel() < wsb:
        ws = _c2s_ws[key] = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    nv = x.shape[0]
    # x + pos reads, bool mask bytes (layers 2-3), the three [nq,128] query-side matrices
    nbytes = 4 * nv * 128 * 2 + (nq * nv if label is not None else 0) + 4 * nq * 128 * 3
    if split:
        with _Timed("c2s", nbytes, 2 * 2 * nv * 128 * heads * nq):
            check(lib().ag3d_c2s_attn_fwd_split(_p(x), _p(pos), nv, _p(qfold), nq, heads, _p(label), _p(q_obj), _p(obj_count),
                                                _p(ctx), _p(lse), _p(ws), ws.numel(), _stream()), "ag3d_c2s_attn_fwd_split")
        return ctx
    with _Timed("c2s", nbytes, 2 * 2 * nv * 128 * heads * nq):
        check(lib().ag3d_c2s_attn_fwd(_p(x), _p(pos), nv, _p(qfold), nq, heads, _p(label), _p(q_obj),
                                      _p(obj_count), _p(ctx), _p(lse), algo, _p(ws), ws.numel(), _stream()),
              "ag3d_c2s_attn_fwd")
    return ctx


S2C_MAX_QUERIES = 256       # per scene: 10 learned background queries + clicks (eval_multi_obj.py:116-167 reaches 210)


S2C_SPLIT_MAX_QUERIES = 24  # the split-row (TMA-fed) variant: heads padded to 16 or 24 query columns


def s2c_mask_fwd(x, pos, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out=None, algo=ALGO_AUTO,
                 split=False, write_x=True):
    """-> (x_out f32 [Nv,128], logits f32 [Nv,n_obj], label u8 [Nv], obj_count i32 [n_obj]).
    split=True: x, pos and x_out are "split" rows (nq <= 24); write_x=False skips the feature write (last layer)."""
    _need_cuda(x, pos, A, c, U, E, q_obj)
    for t in (x, pos, A, c, U, bo, ln_w, ln_b, E):
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise _lib.Ag3dError("s2c inputs must be contiguous fp32")
    if nq > S2C_MAX_QUERIES:
        raise _lib.Ag3dError(f"at most {S2C_MAX_QUERIES} click queries per scene, got {nq}")
    nv = x.shape[0]
    if x_out is None and (write_x or not split):
        x_out = torch.empty_like(x)
    logits = torch.empty((nv, n_obj), dtype=torch.float32, device=x.device)
    label = torch.empty(nv, dtype=torch.uint8, device=x.device)
    obj_count = torch.zeros(n_obj, dtype=torch.int32, device=x.device)
    # s2c (x, pos reads + x' write) + mask head (x' read, logits write, [nq,nv] bool mask write)
    nbytes = 4 * nv * 128 * 3 + 4 * nv * 128 + 4 * nv * n_obj + nq * nv
    ws, wsb = None, 0
    if algo != ALGO_SIMT:
        wsb = lib().ag3d_s2c_workspace_bytes(nq)
        ws = _workspace("s2c", x.device, wsb) if wsb else None
    if split:
        if nq > S2C_SPLIT_MAX_QUERIES or algo == ALGO_SIMT:
            raise _lib.Ag3dError(f"the split-row s2c kernel handles at most {S2C_SPLIT_MAX_QUERIES} queries on the tensor-core path")
        xo = x_out if write_x else None
        with _Timed("s2c_mask", nbytes, 2 * nv * 128 * (2 * heads * nq + nq)):
            check(lib().ag3d_s2c_mask_fwd_split(_p(x), _p(pos), nv, _p(A), _p(c), _p(U), _p(bo), _p(ln_w), _p(ln_b),
                                                float(ln_eps), _p(E), _p(q_obj), nq, heads, n_obj, _p(xo), _p(logits),
                                                _p(label), _p(obj_count), _p(ws), wsb, _stream()), "ag3d_s2c_mask_fwd_split")
        return xo, logits, label, obj_count
    with _Timed("s2c_mask", nbytes, 2 * nv * 128 * (2 * heads * nq + nq)):
        check(lib().ag3d_s2c_mask_fwd(_p(x), _p(pos), nv, _p(A), _p(c), _p(U), _p(bo), _p(ln_w), _p(ln_b),
                                      float(ln_eps), _p(E), _p(q_obj), nq, heads, n_obj, _p(x_out), _p(logits),
                                      _p(label), _p(obj_count), algo, _p(ws), wsb, _stream()), "ag3d_s2c_mask_fwd")
    return x_out, logits, label, obj_count


# =============================================================================================== interactive loop + voxelisation
def click_pred(logits, nv, n_obj, click_rows=None, click_objs=None):
    """pred int32 [nv] = argmax of logits (None: zeros) with the clicked voxels overwritten (eval_multi_obj.py:124-139)."""
    dev = click_rows.device if logits is None else logits.device
    pred = torch.empty(nv, dtype=torch.int32, device=dev)
    n_clicks = 0 if click_rows is None else int(click_rows.shape[0])
    check(lib().ag3d_click_pred(_p(logits), n_obj, nv, _p(click_rows), _p(click_objs), n_clicks, _p(pred), _stream()),
          "ag3d_click_pred")
    return pred


def scene_iou_counts(pred, inverse_map, labels_full, n_obj):
    """-> uint64-as-int64 [n_obj, 3] = (intersection, |pred == o|, |label == o|) at full resolution."""
    _need_cuda(pred, labels_full)
    counts = torch.empty((n_obj, 3), dtype=torch.int64, device=pred.device)
    check(lib().ag3d_scene_iou(_p(pred), _p(inverse_map), _p(labels_full), labels_full.shape[0], n_obj, _p(counts), _stream()),
          "ag3d_scene_iou")
    return counts


def click_simulate(pred, gt, xyz, top_n=-1, perm=None, max_new=32):
    """utils/seg.py:173-226 on the device -> int32 [1 + 4 max_new] (see include/agile3d_b200.h)."""
    _need_cuda(pred, gt, xyz)
    nv = pred.shape[0]
    out = torch.zeros(1 + 4 * max_new, dtype=torch.int32, device=pred.device)
    wsb = lib().ag3d_click_simulate_workspace_bytes(nv)
    ws = _workspace("click", pred.device, wsb)
    check(lib().ag3d_click_simulate(_p(pred), _p(gt), _p(xyz), nv, top_n, _p(perm), max_new, _p(out), _p(ws), ws.numel(),
                                    _stream()), "ag3d_click_simulate")
    return out


def quantize_unique(points, quantization_size, batch_index=0):
    """ME.utils.sparse_quantize on the device: points f32 [n,3] -> (coords int32 [m,4] unique voxels in first-occurrence
    order, unique_map int64 [m], inverse_map int64 [n]).  One host read-back (m)."""
    _need_cuda(points)
    pts = points.contiguous().float()
    n, dev = pts.shape[0], pts.device
    coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    check(lib().ag3d_quantize_points(_p(pts), n, float(quantization_size), batch_index, _p(coords), _p(status), _stream()),
          "ag3d_quantize_points")
    cap = lib().ag3d_hash_capacity(n)
    table = torch.empty(cap * SLOT_BYTES, dtype=torch.uint8, device=dev)
    parent = torch.empty(n, dtype=torch.int32, device=dev)
    out = torch.empty((n, 4), dtype=torch.int32, device=dev)
    wsb = lib().ag3d_downsample_workspace_bytes(n)
    ws = _workspace("downsample", dev, wsb)
    check(lib().ag3d_downsample(_p(coords), n, 1, _p(table), cap, _p(parent), _p(out), _p(status[1:]), _p(ws), ws.numel(),
                                _stream()), "ag3d_downsample")
    bad, m = status.tolist()
    if bad:
        raise ValueError(f"{bad} points quantise outside +-32767 voxels")
    unique_map = torch.empty(m, dtype=torch.int64, device=dev)
    check(lib().ag3d_first_rows(_p(parent), n, m, _p(unique_map), _stream()), "ag3d_first_rows")
    return out[:m], unique_map, parent.to(torch.int64)


# =============================================================================================== click-query side (K11)
def query_blob_floats():
    return int(lib().ag3d_query_blob_floats())


def query_init(feats, xyz, rng, src_row, time_idx, scene_of_row, gauss_B, time_table, bg_feat, bg_pos, feat_row=None,
               feats_split=False):
    """-> (queries, qpos) f32 [rows, 128]; src_row int32 [rows] (>= 0: clicked voxel row of xyz, -(k+1): learned bg query k);
    feat_row (optional): the clicked voxels' rows in `feats` when it is stored in another row order; feats_split: `feats`
    holds "split" rows."""
    _need_cuda(feats, xyz, rng, src_row)
    n = src_row.shape[0]
    q = torch.empty((n, 128), dtype=torch.float32, device=feats.device)
    qp = torch.empty_like(q)
    check(lib().ag3d_query_init(_p(feats), _p(xyz), _p(rng), _p(src_row), _p(feat_row), _p(time_idx), _p(scene_of_row), n, _p(gauss_B),
                                _p(time_table), _p(bg_feat), _p(bg_pos), _p(q), _p(qp), 1 if feats_split else 0, _stream()),
          "ag3d_query_init")
    return q, qp


def query_fold_c2s(queries, qpos, blob, B, nq, heads=8):
    _need_cuda(queries, qpos, blob)
    qfold = torch.empty((B, heads * nq, 128), dtype=torch.float32, device=queries.device)
    check(lib().ag3d_query_fold_c2s(_p(queries), _p(qpos), _p(blob), B, nq, _p(qfold), _stream()), "ag3d_query_fold_c2s")
    return qfold


def query_update_a(ctx, queries, qpos, blob, B, nq, ln_eps=1e-5):
    """tail of c2s + the c2c projections -> (q1, qh, kh, vh), each f32 [B, nq, 128]."""
    _need_cuda(ctx, queries, qpos, blob)
    out = torch.empty((4, B, nq, 128), dtype=torch.float32, device=queries.device)
    check(lib().ag3d_query_update_a(_p(ctx), _p(queries), _p(qpos), _p(blob), B, nq, float(ln_eps), _p(out[0]), _p(out[1]),
                                    _p(out[2]), _p(out[3]), _stream()), "ag3d_query_update_a")
    return out[0], out[1], out[2], out[3]


def query_update_b(q1, qh, kh, vh, qpos, blob, B, nq, heads=8, ln_eps=1e-5):
    """c2c attention + FFN + s2c folds + mask embeddings -> (q3 [B,nq,128], A [B,8nq,128], c [B,8nq], U [B,8nq,128], E [B,nq,128])."""
    _need_cuda(q1, qh, kh, vh, qpos, blob)
    dev, f32 = q1.device, torch.float32
    q3 = torch.empty((B, nq, 128), dtype=f32, device=dev)
    A = torch.empty((B, heads * nq, 128), dtype=f32, device=dev)
    c = torch.empty((B, heads * nq), dtype=f32, device=dev)
    U = torch.empty((B, heads * nq, 128), dtype=f32, device=dev)
    E = torch.empty((B, nq, 128), dtype=f32, device=dev)
    check(lib().ag3d_query_update_b(_p(q1), _p(qh), _p(kh), _p(vh), _p(qpos), _p(blob), B, nq, float(ln_eps), _p(q3), _p(A),
                                    _p(c), _p(U), _p(E), _stream()), "ag3d_query_update_b")
    return q3, A, c, U, E


# =============================================================================================== training step
def _ws_for(tag, device, nbytes):
    return _workspace(tag, device, max(int(nbytes), 16))


def bn_stats(z, eps, momentum=0.0, running_mean=None, running_var=None):
    """batch mean / inverse std of the rows of z [N,C] (+ running-stat update) -> (mean [C], invstd [C])."""
    _need_cuda(z)
    zp, z_ld = _rows2d(z)
    n, c = z.shape
    mean = torch.empty(c, dtype=torch.float32, device=z.device)
    invstd = torch.empty(c, dtype=torch.float32, device=z.device)
    wsb = lib().ag3d_colreduce_workspace_bytes(c)
    ws = _ws_for("colreduce", z.device, wsb)
    with _Timed("bn", 4 * n * c):
        check(lib().ag3d_bn_stats(zp, z_ld, c, n, float(eps), float(momentum), _p(running_mean), _p(running_var),
                                  _p(mean), _p(invstd), _p(ws), ws.numel(), _stream()), "ag3d_bn_stats")
    return mean, invstd


def bn_apply(z, mean, invstd, gamma, beta, out, residual=None, relu=False):
    _need_cuda(z, out)
    zp, z_ld = _rows2d(z)
    op, o_ld = _rows2d(out)
    rp, r_ld = (_rows2d(residual) if residual is not None else (C.c_void_p(0), 0))
    n, c = z.shape
    with _Timed("bn", 4 * n * c * (3 if residual is not None else 2)):
        check(lib().ag3d_bn_apply(zp, z_ld, _p(mean), _p(invstd), _p(gamma), _p(beta), rp, r_ld, c, n,
                                  RELU if relu else 0, op, o_ld, _stream()), "ag3d_bn_apply")
    return out


def bn_bwd(z, y, dy, mean, invstd, gamma, dz, relu=False, g_out=None):
    """-> (dgamma [C], dbeta [C]); writes dz (may alias dy) and, when asked, the relu-masked dy into g_out."""
    _need_cuda(z, dy, dz)
    zp, z_ld = _rows2d(z)
    yp, y_ld = (_rows2d(y) if y is not None else (C.c_void_p(0), 0))
    dyp, dy_ld = _rows2d(dy)
    dzp, dz_ld = _rows2d(dz)
    gp, g_ld = (_rows2d(g_out) if g_out is not None else (C.c_void_p(0), 0))
    n, c = z.shape
    dgamma = torch.empty(c, dtype=torch.float32, device=z.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=z.device)
    wsb = lib().ag3d_colreduce_workspace_bytes(c)
    ws = _ws_for("colreduce", z.device, wsb)
    with _Timed("bn", 4 * n * c * 6):
        check(lib().ag3d_bn_bwd(zp, z_ld, yp, y_ld, dyp, dy_ld, _p(mean), _p(invstd), _p(gamma), c, n,
                                RELU if relu else 0, dzp, dz_ld, gp, g_ld, _p(dgamma), _p(dbeta), _p(ws), ws.numel(),
                                _stream()), "ag3d_bn_bwd")
    return dgamma, dbeta


def col_sum(z):
    _need_cuda(z)
    zp, z_ld = _rows2d(z)
    n, c = z.shape
    out = torch.empty(c, dtype=torch.float32, device=z.device)
    wsb = lib().ag3d_colreduce_workspace_bytes(c)
    ws = _ws_for("colreduce", z.device, wsb)
    check(lib().ag3d_col_sum(zp, z_ld, c, n, _p(out), _p(ws), ws.numel(), _stream()), "ag3d_col_sum")
    return out


def pack_split_rows(x):
    """fp32 rows [N,C] (C % 32 == 0, a channel slice is fine) -> new contiguous tensor of "split" rows (one kernel)."""
    _need_cuda(x)
    xp, x_ld = _rows2d(x)
    n, c = x.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    check(lib().ag3d_pack_split(xp, x_ld, c, n, _p(out), c, _stream()), "ag3d_pack_split")
    return out


def wgrad_tc_supported(K, cin, cout):
    return bool(lib().ag3d_spconv_bwd_weight_tc_supported(K, cin, cout))


def spconv_bwd_weight_tc(x_split, nbr, dout_split, K, dweight=None, accumulate=False):
    """Tensor-core weight gradient: both operands are split rows (pack_split_rows).  -> f32 [K, cin, cout]."""
    _need_cuda(x_split, dout_split)
    xp, x_ld = _rows2d(x_split)
    dp, d_ld = _rows2d(dout_split)
    n_out, cout = dout_split.shape
    n_in, cin = x_split.shape
    if nbr is not None and (tuple(nbr.shape) != (K, n_out) or not nbr.is_contiguous()):
        raise _lib.Ag3dError(f"neighbour table must be contiguous int32 [{K},{n_out}], got {tuple(nbr.shape)}")
    if nbr is None and (K != 1 or n_in != n_out):
        raise _lib.Ag3dError("identity-map weight gradient needs K == 1 and equal row counts")
    if dweight is None:
        dweight = torch.empty((K, cin, cout), dtype=torch.float32, device=x_split.device)
        accumulate = False
    wsb = lib().ag3d_spconv_bwd_weight_tc_workspace_bytes(n_out, K, cin, cout)
    ws = _ws_for("wgrad_tc", x_split.device, wsb)
    pairs = n_out
    if _prof is not None and nbr is not None:
        pairs = int((nbr >= 0).sum().item())
    with _Timed("spconv_wgrad", 4 * n_in * cin + 4 * n_out * cout + 4 * K * cin * cout + 4 * pairs, 2 * pairs * cin * cout):
        check(lib().ag3d_spconv_bwd_weight_tc(xp, n_in, x_ld, cin, _p(nbr), K, n_out, dp, d_ld, cout, _p(dweight),
                                              1 if accumulate else 0, _p(ws), ws.numel(), _stream()),
              "ag3d_spconv_bwd_weight_tc")
    return dweight


def spconv_bwd_weight(x, nbr, dout, K, dweight=None, accumulate=False):
    """dW[k] (+)= x[nbr[k]]^T dout  -> f32 [K, cin, cout].  nbr None: K = 1 on the identity map (plain X^T dY)."""
    _need_cuda(x, dout)
    xp, x_ld = _rows2d(x)
    dp, d_ld = _rows2d(dout)
    n_out, cout = dout.shape
    cin = x.shape[1]
    if nbr is not None and (tuple(nbr.shape) != (K, n_out) or not nbr.is_contiguous()):
        raise _lib.Ag3dError(f"neighbour table must be contiguous int32 [{K},{n_out}], got {tuple(nbr.shape)}")
    if nbr is None and (K != 1 or x.shape[0] != n_out):
        raise _lib.Ag3dError("identity-map weight gradient needs K == 1 and equal row counts")
    if dweight is None:
        dweight = torch.empty((K, cin, cout), dtype=torch.float32, device=x.device)
        accumulate = False
    if dweight.numel() != K * cin * cout or not dweight.is_contiguous():
        raise _lib.Ag3dError("dweight must be contiguous [K,cin,cout]")
    wsb = lib().ag3d_spconv_bwd_weight_workspace_bytes(n_out, K, cin, cout)
    ws = _ws_for("wgrad", x.device, wsb)
    with _Timed("spconv_wgrad", 4 * x.shape[0] * cin + 4 * n_out * cout + 4 * K * cin * cout, 0):
        check(lib().ag3d_spconv_bwd_weight(xp, x_ld, cin, _p(nbr), K, n_out, dp, d_ld, cout, _p(dweight),
                                           1 if accumulate else 0, _p(ws), ws.numel(), _stream()),
              "ag3d_spconv_bwd_weight")
    return dweight


def stem_bwd_weight(coords, feats, table, cap, ksize, dz, bricks=None):
    _need_cuda(coords, feats, dz)
    dp, d_ld = _rows2d(dz)
    n = coords.shape[0]
    dw = torch.empty((ksize ** 3, 3, 32), dtype=torch.float32, device=dz.device)
    wsb = lib().ag3d_stem_bwd_weight_workspace_bytes(ksize)
    ws = _ws_for("stem_wgrad", dz.device, wsb)
    with _Timed("stem_wgrad", 16 * n + 12 * n + 4 * n * 32 + 16 * ksize ** 3 * n):
        if bricks is not None and ksize in (3, 5):
            t2, cap2, rows = bricks
            check(lib().ag3d_stem_bwd_weight_bricks(_p(coords), _p(feats), n, _p(t2), cap2, _p(rows), ksize, dp, d_ld,
                                                    _p(dw), 0, _p(ws), ws.numel(), _stream()), "ag3d_stem_bwd_weight_bricks")
        else:
            check(lib().ag3d_stem_bwd_weight(_p(coords), _p(feats), n, _p(table), cap, ksize, dp, d_ld, _p(dw), 0, _p(ws),
                                             ws.numel(), _stream()), "ag3d_stem_bwd_weight")
    return dw


def decoder_bwd_rows(nq, heads):
    r = int(lib().ag3d_decoder_bwd_rows(nq, heads))
    if r == 0:
        raise _lib.Ag3dError("decoder backward handles at most 32 click queries per scene")
    return r


def c2s_attn_bwd(x, pos, qf, qft, dctx, dctxt, lse, dr, rowobj, hqp, label):
    """-> (dx [Nv,128], dS [Nv,hqp])."""
    _need_cuda(x, pos, qf)
    nv = x.shape[0]
    dx = torch.empty_like(x)
    ds = torch.empty((nv, hqp), dtype=torch.float32, device=x.device)
    with _Timed("c2s_bwd", 4 * nv * 128 * 3 + 4 * nv * hqp, 2 * 4 * nv * 128 * hqp):
        check(lib().ag3d_c2s_attn_bwd(_p(x), _p(pos), nv, _p(qf), _p(qft), _p(dctx), _p(dctxt), _p(lse), _p(dr),
                                      _p(rowobj), hqp, _p(label), _p(dx), _p(ds), _stream()), "ag3d_c2s_attn_bwd")
    return dx, ds


def _pad32(n):
    return (n + 31) // 32 * 32


def _pad_rows(t, rows):
    if t.shape[0] == rows:
        return t.contiguous()
    out = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


def c2s_attn_bwd_tc(x, pos, qf, dctx, lse, dr, rowobj, hqp, label):
    """Backward of the click -> scene attention with its four GEMMs as 1x1 tensor-core convolutions (bf16x3) and one
    point-wise kernel in between, in chunks of at most 256 (head, query) rows (any number of click queries):
        S = (x+pos) qf^T,  dP = x dctx^T,  P = masked exp(S - lse),  dS = P (dP - dr),  dx = P dctx + dS qf.
    qf, dctx: [hqp, 128]; lse, dr, rowobj: [hqp] (rows padded to hqp, a multiple of 32).
    -> (dx [Nv,128], [dS chunk [Nv, w_j]] in row order)."""
    _need_cuda(x, pos, qf, dctx)
    nv, d = x.shape
    f32 = dict(dtype=torch.float32, device=x.device)
    xp = x + pos
    dx, ds_chunks = None, []
    for a in range(0, hqp, 256):
        b = min(a + 256, hqp)
        w = b - a
        qf_j, dc_j = qf[a:b].contiguous(), dctx[a:b].contiguous()
        qft, dct = qf_j.t().contiguous(), dc_j.t().contiguous()
        s_p, dp_ds = torch.empty((nv, w), **f32), torch.empty((nv, w), **f32)
        spconv_fwd(xp, None, qft, s_p, algo=ALGO_TC, weight_tc=prepare_tc_weight(qft))
        spconv_fwd(x, None, dct, dp_ds, algo=ALGO_TC, weight_tc=prepare_tc_weight(dct))
        with _Timed("c2s_bwd", 4 * nv * w * 4):
            check(lib().ag3d_c2s_bwd_pointwise(_p(s_p), _p(dp_ds), _p(lse[a:b]), _p(dr[a:b]), _p(rowobj[a:b]), _p(label), nv, w,
                                               _stream()), "ag3d_c2s_bwd_pointwise")
        dx0 = torch.empty((nv, d), **f32)
        spconv_fwd(s_p, None, dc_j, dx0, residual=dx, algo=ALGO_TC, weight_tc=prepare_tc_weight(dc_j))
        dx = torch.empty((nv, d), **f32)
        spconv_fwd(dp_ds, None, qf_j, dx, residual=dx0, algo=ALGO_TC, weight_tc=prepare_tc_weight(qf_j))
        ds_chunks.append(dp_ds)
    return dx, ds_chunks


def s2c_mask_bwd(x, pos, A, At, c, U, Ut, bo, ln_w, ln_b, ln_eps, E, Et, q_obj, nq, heads, n_obj, hqp, dxo, dlogits):
    """-> (dx, a [Nv,hqp], dS [Nv,hqp], dy [Nv,128], g [Nv,32], colsums [3*128+hqp])."""
    _need_cuda(x, pos, A)
    nv, dev = x.shape[0], x.device
    f32 = dict(dtype=torch.float32, device=dev)
    dx = torch.empty_like(x)
    a = torch.empty((nv, hqp), **f32)
    ds = torch.empty((nv, hqp), **f32)
    dy = torch.empty((nv, 128), **f32)
    g = torch.empty((nv, 32), **f32)
    cols = torch.empty(3 * 128 + hqp, **f32)
    wsb = lib().ag3d_s2c_bwd_workspace_bytes(hqp)
    ws = _ws_for("s2c_bwd", dev, wsb)
    with _Timed("s2c_bwd", 4 * nv * 128 * 5 + 8 * nv * hqp, 2 * nv * 128 * (4 * hqp + 64)):
        check(lib().ag3d_s2c_mask_bwd(_p(x), _p(pos), nv, _p(A), _p(At), _p(c), _p(U), _p(Ut), _p(bo), _p(ln_w),
                                      _p(ln_b), float(ln_eps), _p(E), _p(Et), _p(q_obj), nq, heads, n_obj, hqp,
                                      _p(dxo), _p(dlogits), _p(dx), _p(a), _p(ds), _p(dy), _p(g), _p(cols), _p(ws),
                                      ws.numel(), _stream()), "ag3d_s2c_mask_bwd")
    return dx, a, ds, dy, g, cols


def s2c_mask_bwd_tc_any(x, pos, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, dxo, dlogits, x_out, xt_dy):
    """Backward of ag3d_s2c_mask_fwd for ANY number of click queries (<= 256), composed of 1x1 tensor-core convolutions
    and row-wise kernels; the (head, query) columns are processed in chunks of whole heads of at most 256 columns:
        S = (x+pos) A^T + c -> a = per-head softmax        y = x + a U + bo -> n, rstd (LayerNorm statistics)
        G = x' E^T -> g = first-maximum routing of dlogits  t = g E + dxo -> dy = LayerNorm backward (+ column sums)
        da = dy U^T -> dS = a (da - <a, da>)                dx = dy + dS A
    A, U: [heads*nq, 128] (rows h-major), c: [heads*nq], E: [nq, 128]; x_out = x' saved by the forward;
    xt_dy(list of [Nv,cin], [Nv,cout]) -> [1,cin,cout] is the caller's X^T dY contraction.
    -> (dx, dA, dc, dU, dbo, dln_w, dln_b, dE)."""
    _need_cuda(x, pos, A, U, E)
    nv, d = x.shape
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    st = _stream()
    conv = lambda inp, w, out, **kw: spconv_fwd(inp, None, w, out, algo=ALGO_TC, weight_tc=prepare_tc_weight(w), **kw)
    xp = x + pos
    hc = max(1, 256 // nq)                                            # whole heads per chunk
    chunks = [(h0, min(h0 + hc, heads)) for h0 in range(0, heads, hc)]
    parts, y = [], torch.empty((nv, d), **f32)
    for j, (h0, h1) in enumerate(chunks):
        rows = slice(h0 * nq, h1 * nq)
        w = (h1 - h0) * nq
        wp = _pad32(w)
        A_j, U_j, c_j = _pad_rows(A[rows], wp), _pad_rows(U[rows], wp), _pad_rows(c[rows], wp)
        a = torch.empty((nv, wp), **f32)
        conv(xp, A_j.t().contiguous(), a, shift=c_j)
        with _Timed("s2c_bwd", 8 * nv * wp):
            check(lib().ag3d_s2c_softmax_heads(_p(a), nv, h1 - h0, nq, wp, st), "ag3d_s2c_softmax_heads")
        if j == 0:
            conv(a, U_j, y, shift=bo, residual=x)
        else:
            conv(a, U_j, y, residual=y)
        parts.append((rows, w, wp, A_j, U_j, a))
    n, rstd = torch.empty((nv, d), **f32), torch.empty(nv, **f32)
    with _Timed("s2c_bwd", 8 * nv * d):
        check(lib().ag3d_ln_fwd_stats(_p(y), nv, float(ln_eps), _p(n), _p(rstd), st), "ag3d_ln_fwd_stats")
    nqp = _pad32(nq)
    E_p = _pad_rows(E, nqp)
    G, g = torch.empty((nv, nqp), **f32), torch.empty((nv, nqp), **f32)
    conv(x_out, E_p.t().contiguous(), G)
    arg = torch.empty((nv, n_obj), dtype=torch.int32, device=dev)
    check(lib().ag3d_s2c_route_ld(_p(G), nqp, _p(dlogits), _p(q_obj), nq, n_obj, nv, _p(arg), _p(g), st), "ag3d_s2c_route_ld")
    t = y                                                             # reuse: y is no longer needed
    if dxo is not None:
        conv(g, E_p, t, residual=dxo)
    else:
        conv(g, E_p, t)
    dy = torch.empty((nv, d), **f32)
    cols = torch.empty(3 * d + 256, **f32)
    wsb = lib().ag3d_ln_bwd_workspace_bytes()
    ws = _ws_for("ln_bwd", dev, wsb)
    with _Timed("s2c_bwd", 16 * nv * d):
        check(lib().ag3d_ln_bwd(_p(t), _p(n), _p(rstd), _p(ln_w), nv, _p(dy), _p(cols), _p(ws), ws.numel(), st), "ag3d_ln_bwd")
    hq = heads * nq
    dA, dU, dc = torch.empty((hq, d), **f32), torch.empty((hq, d), **f32), torch.empty(hq, **f32)
    dx = None
    for rows, w, wp, A_j, U_j, a in parts:
        ds = torch.empty((nv, wp), **f32)
        conv(dy, U_j.t().contiguous(), ds)
        with _Timed("s2c_bwd", 12 * nv * wp):
            check(lib().ag3d_s2c_ds(_p(a), _p(ds), nv, w // nq, nq, wp, st), "ag3d_s2c_ds")
        dc[rows] = col_sum(ds)[:w]
        dx_new = torch.empty((nv, d), **f32)
        conv(ds, A_j, dx_new, residual=dy if dx is None else dx)
        dx = dx_new
        dA[rows] = xt_dy([ds], xp)[0, :w]
        dU[rows] = xt_dy([a], dy)[0, :w]
    dE = xt_dy([g], x_out)[0, :nq]
    return dx, dA, dc, dU, cols[:d], cols[d:2 * d], cols[2 * d:3 * d], dE


def loss_fwd(logits, target, w, eps=1e-6):
    """-> f32 [2] = (sum_v w ce_v, sum_v w dice_v) for one scene."""
    _need_cuda(logits, target, w)
    n, c = logits.shape
    sums = torch.empty(4, dtype=torch.float32, device=logits.device)
    wsb = lib().ag3d_loss_workspace_bytes()
    ws = _ws_for("loss", logits.device, wsb)
    check(lib().ag3d_loss_fwd(_p(logits), c, n, _p(target), _p(w), float(eps), _p(sums), _p(ws), ws.numel(), _stream()),
          "ag3d_loss_fwd")
    return sums[0::2]


def loss_bwd(logits, target, w, g, eps=1e-6):
    n, c = logits.shape
    d = torch.empty_like(logits)
    check(lib().ag3d_loss_bwd(_p(logits), c, n, _p(target), _p(w), float(eps), _p(g), _p(d), _stream()), "ag3d_loss_bwd")
    return d


def click_loss_weights(xyz, clicks, alpha=0.8, beta=2.0, tita=0.3):
    _need_cuda(xyz, clicks)
    n = xyz.shape[0]
    w = torch.empty(n, dtype=torch.float32, device=xyz.device)
    check(lib().ag3d_click_loss_weights(_p(xyz), n, _p(clicks), clicks.shape[0], alpha, beta, tita, _p(w), _stream()),
          "ag3d_click_loss_weights")
    return w


def grad_norm(flat):
    _need_cuda(flat)
    out = torch.empty(1, dtype=torch.float32, device=flat.device)
    wsb = lib().ag3d_grad_norm_workspace_bytes()
    ws = _ws_for("gnorm", flat.device, wsb)
    check(lib().ag3d_grad_norm(_p(flat), flat.numel(), _p(out), _p(ws), ws.numel(), _stream()), "ag3d_grad_norm")
    return out


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, norm=None, max_norm=0.0):
    _need_cuda(p, g, m, v)
    check(lib().ag3d_adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                _p(norm), float(max_norm), _stream()), "ag3d_adamw_step")
