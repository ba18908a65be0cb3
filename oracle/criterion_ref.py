"""CPU restatement of the reference's training loss.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU leg).

Restates (own code; pinned against the unmodified reference files by tests/golden/make_golden.py, which imports
models/criterion.py and utils/seg.py by path and stores their outputs in tests/golden/train_*.npz):
  * SetCriterion.forward / loss_bce / loss_dice / multiclass_dice_loss / dice_loss   models/criterion.py:15-132
  * build_mask_criterion                                                              models/criterion.py:135-153
  * loss_weights / cal_click_loss_weights                                             utils/seg.py:62-89
  * the scalar the trainer back-propagates: sum_k loss_dict[k] * weight_dict[k]       engine.py:126-128
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dice_per_voxel(logits, target, eps=1e-6):
    """criterion.py:15-75: softmax over classes, one-hot target, then dice over the CLASS axis of every voxel."""
    p = logits.softmax(1)
    onehot = F.one_hot(target.long(), logits.shape[1]).to(p.dtype)
    num = 2.0 * (p * onehot).mean(1)
    den = (p + onehot).mean(1)
    iou = (num + eps) / (den + eps)
    return torch.where(num > eps, 1.0 - iou, iou * 0.0)


def criterion(outputs, targets, weights, losses=("bce", "dice")):
    """-> {'loss_bce', 'loss_dice', 'loss_bce_0', ...} exactly as SetCriterion.forward (criterion.py:114-132)."""
    out = {}

    def emit(pred, suffix):
        masks = pred["pred_masks"]
        if "bce" in losses:
            out["loss_bce" + suffix] = sum((F.cross_entropy(m, t.long(), reduction="none") * w).mean()
                                           for m, t, w in zip(masks, targets, weights)) / len(masks)
        if "dice" in losses:
            out["loss_dice" + suffix] = sum((dice_per_voxel(m, t) * w).mean()
                                            for m, t, w in zip(masks, targets, weights)) / len(masks)

    emit(outputs, "")
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        emit(aux, f"_{i}")
    return out


def weight_dict(args):
    """criterion.py:135-147."""
    wd = {"loss_bce": args.bce_loss_coef, "loss_dice": args.dice_loss_coef}
    if args.aux:
        for i in range(args.num_decoders * len(args.hlevels)):
            wd[f"loss_bce_{i}"] = args.bce_loss_coef
            wd[f"loss_dice_{i}"] = args.dice_loss_coef
    return wd


def total_loss(loss_dict, wd):
    """engine.py:128."""
    return sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)


def click_loss_weights(xyz, click_rows, alpha=0.8, beta=2.0, tita=0.3):
    """utils/seg.py:62-70 for one scene: xyz [Nv,3], click_rows list[int] -> [Nv]."""
    d = torch.cdist(xyz, xyz[click_rows]).min(dim=1)[0]
    return alpha + (beta - alpha) * (1 - torch.clamp(d, max=tita) / tita)
