"""ctypes binding of csrc/libagile3d_b200.so (C-ABI: include/agile3d_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C agile3d_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libagile3d_b200.so")

RELU = 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_TC_PACKED = 0, 1, 2, 3

_lib = None

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every entry of include/agile3d_b200.h appears here
SIGNATURES = {
    "ag3d_abi_version": (_i32, []),
    "ag3d_last_error": (C.c_char_p, []),
    "ag3d_device_info": (_i32, [_vp, _vp, _vp]),
    "ag3d_kernel_launches": (_i64, []),
    "ag3d_hash_capacity": (_i64, [_i64]),
    "ag3d_hash_build": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "ag3d_downsample_workspace_bytes": (_sz, [_i64]),
    "ag3d_downsample": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_row_order_workspace_bytes": (_sz, [_i64]),
    "ag3d_row_order": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_permute_map": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "ag3d_downsample_dev": (_i32, [_vp, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_scene_offsets": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "ag3d_kernel_map": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ag3d_kernel_map_transposed": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "ag3d_spconv_tc_weight_bytes": (_sz, [_i32, _i32, _i32]),
    "ag3d_spconv_tc_prepare_weight": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "ag3d_spconv_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "ag3d_spconv_fwd": (_i32, [_vp, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _i32,
                               _i32, _vp, _sz, _vp]),
    "ag3d_spconv_fwd_rows": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _i32,
                                    _i32, _i32, _vp, _sz, _vp]),
    "ag3d_stem_conv_fwd": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "ag3d_gather_rows": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "ag3d_brick_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "ag3d_stem_conv_fwd_bricks": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "ag3d_posenc_workspace_bytes": (_sz, [_i32]),
    "ag3d_fourier_posenc": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_fourier_posenc_split": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_c2s_workspace_bytes": (_sz, [_i32, _i32]),
    "ag3d_c2s_attn_fwd": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_c2s_attn_fwd_split": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_s2c_workspace_bytes": (_sz, [_i32]),
    "ag3d_s2c_mask_fwd": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _i32, _i32,
                                 _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_s2c_mask_fwd_split": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _i32, _i32,
                                       _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_click_pred": (_i32, [_vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp]),
    "ag3d_scene_iou": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "ag3d_click_simulate_workspace_bytes": (_sz, [_i64]),
    "ag3d_click_simulate": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "ag3d_quantize_points": (_i32, [_vp, _i64, _f32, _i32, _vp, _vp, _vp]),
    "ag3d_first_rows": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "ag3d_query_blob_floats": (_i64, []),
    "ag3d_query_init": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "ag3d_query_fold_c2s": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "ag3d_query_update_a": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "ag3d_query_update_b": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    # ---- training step
    "ag3d_colreduce_workspace_bytes": (_sz, [_i32]),
    "ag3d_bn_stats": (_i32, [_vp, _i32, _i32, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_bn_apply": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _vp, _i32, _vp]),
    "ag3d_bn_bwd": (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _vp, _i32,
                           _vp, _vp, _vp, _sz, _vp]),
    "ag3d_col_sum": (_i32, [_vp, _i32, _i32, _i64, _vp, _vp, _sz, _vp]),
    "ag3d_spconv_bwd_data": (_i32, [_vp, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp,
                                    _sz, _vp]),
    "ag3d_spconv_bwd_weight_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "ag3d_spconv_bwd_weight": (_i32, [_vp, _i32, _i32, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_pack_split": (_i32, [_vp, _i32, _i32, _i64, _vp, _i32, _vp]),
    "ag3d_spconv_bwd_weight_tc_supported": (_i32, [_i32, _i32, _i32]),
    "ag3d_spconv_bwd_weight_tc_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "ag3d_spconv_bwd_weight_tc": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_stem_bwd_weight_workspace_bytes": (_sz, [_i32]),
    "ag3d_stem_bwd_weight": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_stem_bwd_weight_bricks": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _sz, _vp]),
    "ag3d_decoder_bwd_rows": (_i32, [_i32, _i32]),
    "ag3d_c2s_attn_bwd": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ag3d_c2s_bwd_pointwise": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "ag3d_s2c_softmax_heads": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp]),
    "ag3d_s2c_ds": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ag3d_ln_fwd_stats": (_i32, [_vp, _i64, _f32, _vp, _vp, _vp]),
    "ag3d_ln_bwd_workspace_bytes": (_sz, []),
    "ag3d_ln_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_s2c_route": (_i32, [_vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp]),
    "ag3d_s2c_route_ld": (_i32, [_vp, _i32, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp]),
    "ag3d_s2c_bwd_workspace_bytes": (_sz, [_i32]),
    "ag3d_s2c_mask_bwd": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _i32, _i32,
                                 _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ag3d_loss_workspace_bytes": (_sz, []),
    "ag3d_loss_fwd": (_i32, [_vp, _i32, _i64, _vp, _vp, _f32, _vp, _vp, _sz, _vp]),
    "ag3d_loss_bwd": (_i32, [_vp, _i32, _i64, _vp, _vp, _f32, _vp, _vp, _vp]),
    "ag3d_click_loss_weights": (_i32, [_vp, _i64, _vp, _i32, _f32, _f32, _f32, _vp, _vp]),
    "ag3d_grad_norm_workspace_bytes": (_sz, []),
    "ag3d_grad_norm": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "ag3d_adamw_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _f32, _vp]),
}


class Ag3dError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ag3dError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
                "agile3d_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.ag3d_abi_version() != 7:
            raise Ag3dError("libagile3d_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().ag3d_last_error()
        raise Ag3dError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
