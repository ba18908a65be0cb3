"""agile3d_b200 — Blackwell-native hot path of AGILE3D (backbone + click-query decoder).

Public surface mirrors the reference (models/__init__.py:6-10, MinkowskiEngine symbols its callers use):
    build_model(args), build_criterion(args), SparseTensor, utils.sparse_quantize, utils.batched_coordinates,
    cal_click_loss_weights (utils/seg.py:72-89)
The CUDA library (csrc/libagile3d_b200.so, C-ABI in include/agile3d_b200.h) is loaded lazily on first
use and there is no CPU fallback: using the model without it raises.
"""
from .minkowski import SparseTensor, batched_coordinates, sparse_quantize, utils  # noqa: F401


def build_model(args):
    from .model import build_agile3d
    return build_agile3d(args)


def build_criterion(args):
    """models/__init__.py:9-10."""
    from .criterion import build_mask_criterion
    return build_mask_criterion(args)


def cal_click_loss_weights(*a, **kw):
    from .criterion import cal_click_loss_weights as f
    return f(*a, **kw)


__all__ = ["build_criterion", "cal_click_loss_weights", "build_model", "SparseTensor", "sparse_quantize", "batched_coordinates", "utils"]
