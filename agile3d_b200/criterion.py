"""SetCriterion (models/criterion.py:7-132) and cal_click_loss_weights (utils/seg.py:62-89) on the B200 library.

Both losses are per-voxel in the reference (the dice term of multiclass_dice_loss reduces over the CLASS axis of
every voxel, criterion.py:70-75), so one kernel pass per prediction computes both sums; autograd calls
ag3d_loss_bwd for d(logits).  Keys, coefficients and the aux-output suffixes follow build_mask_criterion
(criterion.py:135-153).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _LossFn(torch.autograd.Function):
    """logits [Nv, 1+K] f32, target [Nv] int32, w [Nv] f32 -> [2] = (mean_v w ce_v, mean_v w dice_v)."""

    @staticmethod
    def forward(ctx, logits, target, w):
        logits = logits.contiguous()
        ctx.save_for_backward(logits, target, w)
        return ops.loss_fwd(logits, target, w) / logits.shape[0]

    @staticmethod
    def backward(ctx, g):
        logits, target, w = ctx.saved_tensors
        return ops.loss_bwd(logits, target, w, g.contiguous().float()), None, None


class SetCriterion(nn.Module):
    def __init__(self, weight_dict, losses):
        super().__init__()
        self.weight_dict = weight_dict
        self.losses = list(losses)
        for l in self.losses:
            if l not in ("bce", "dice"):
                raise AssertionError(f"do you really want to compute {l} loss?")      # criterion.py:109

    def _pair(self, pred_masks, targets, weights):
        tot = 0.0
        for logits, t, w in zip(pred_masks, targets, weights):
            if t.dtype != torch.int32:
                t = t.to(torch.int32)
            tot = tot + _LossFn.apply(logits, t.contiguous(), w.contiguous().float())
        return tot / len(pred_masks)

    def forward(self, outputs, targets, weights=None):
        if weights is None:
            raise ValueError("SetCriterion needs the click loss weights (utils/seg.py:72-89), as engine.py:124 passes")
        losses = {}

        def emit(pred, suffix):
            both = self._pair(pred["pred_masks"], targets, weights)
            if "bce" in self.losses:
                losses["loss_bce" + suffix] = both[0]
            if "dice" in self.losses:
                losses["loss_dice" + suffix] = both[1]

        emit(outputs, "")
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            emit(aux, f"_{i}")
        return losses


def build_mask_criterion(args):
    """models/criterion.py:135-153."""
    weight_dict = {"loss_bce": args.bce_loss_coef, "loss_dice": args.dice_loss_coef}
    if args.aux:
        aux = {}
        for i in range(args.num_decoders * len(args.hlevels)):
            aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    return SetCriterion(weight_dict, args.losses)


def cal_click_loss_weights(batch_idx, raw_coords, labels, click_idx, alpha=0.8, beta=2.0, tita=0.3):
    """utils/seg.py:72-89: per scene, w_v = alpha + (beta - alpha) * (1 - min(d_v, tita) / tita) with d_v the distance
    from voxel v to the nearest click of the scene.  -> list of [Nv_b] tensors."""
    counts = torch.bincount(batch_idx.long()).tolist()
    raw = raw_coords.float().contiguous()
    out, start = [], 0
    for b, n in enumerate(counts):
        xyz = raw[start:start + n]
        ids = [int(i) for _, v in click_idx[b].items() for i in v]
        clicks = xyz[torch.tensor(ids, dtype=torch.long, device=raw.device)].contiguous()
        out.append(ops.click_loss_weights(xyz, clicks, alpha, beta, tita))
        start += n
    return out
