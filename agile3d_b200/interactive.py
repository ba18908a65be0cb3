"""Device-side mirror of the interactive-loop helpers of the reference (utils/seg.py), same names and return structure,
so that eval_multi_obj.py:118-167 / engine.py:88-115 can call them unchanged on CUDA tensors:

    mean_iou_scene(pred, labels)                                                  utils/seg.py:44-59
    get_simulated_clicks(pred_qv, labels_qv, coords_qv, current_num_clicks, training)    utils/seg.py:173-226
    extend_clicks(current_clicks, current_clicks_time, new_clicks, new_click_time)       utils/seg.py:229-239
    iou_and_simulated_clicks(...)    the first two fused into one read-back per click round (eval_multi_obj.py:140-160)

The reference computes a torch.cdist matrix per error cluster and synchronises several times per cluster; here one
kernel pass handles all clusters (csrc/click_ops.cu) and the host reads back ONE small record per call (the clicks have
to reach the caller's python dicts: that read-back is the interface, not an implementation detail).
"""
from __future__ import annotations

import random

import torch

from . import ops


def mean_iou_scene(pred, labels, inverse_map=None):
    """-> (mean IoU over the objects present in `labels` (id != 0), {object id: IoU}); pred is indexed through inverse_map
    when given (pred[inverse_map] vs full-resolution labels, eval_multi_obj.py:145-148)."""
    n_obj = int(max(int(labels.max()), int(pred.max())) + 1) if labels.numel() else 1
    counts = ops.scene_iou_counts(pred.to(torch.int32).contiguous(), None if inverse_map is None else inverse_map.contiguous(),
                                  labels.to(torch.int32).contiguous(), n_obj).tolist()
    ious = {}
    for o in range(1, n_obj):
        inter, npred, nlab = counts[o]
        if nlab == 0:
            continue                                     # torch.unique(labels) does not contain it
        ious[o] = inter / (npred + nlab - inter)
    mean = sum(ious.values()) / len(ious) if ious else float("nan")
    return torch.tensor(mean), ious


def get_simulated_clicks(pred_qv, labels_qv, coords_qv, current_num_clicks=None, training=True):
    """Same contract as utils/seg.py:173-226: -> (new_clicks {str(obj): [voxel rows]}, number of new clicks,
    new_click_pos {str(obj): [xyz tensors]}, new_click_time {str(obj): [order]}) or four Nones when nothing is wrong.
    The selected clusters are visited in the order of random.shuffle, exactly as the reference does."""
    pred = pred_qv.to(torch.int32).contiguous()
    gt = labels_qv.to(torch.int32).contiguous()
    xyz = coords_qv.float().contiguous()
    if training:
        top_n = int((torch.unique(labels_qv) != 0).sum())          # num_obj (utils/seg.py:190,210-214)
    else:
        top_n = -1 if current_num_clicks == 0 else 1
    max_new = 64
    rec = ops.click_simulate(pred, gt, xyz, top_n=top_n, perm=None, max_new=max_new).tolist()     # the one read-back
    n = rec[0]
    if n == 0:
        return None, None, None, None
    picks = [(rec[1 + 3 * i], rec[2 + 3 * i]) for i in range(n)]     # (row, object), clusters by size, descending
    random.shuffle(picks)                                            # utils/seg.py:127
    clicks, pos, times = {}, {}, {}
    for order, (row, obj) in enumerate(picks):
        clicks.setdefault(str(obj), []).append(int(row))
        pos.setdefault(str(obj), []).append(coords_qv[row])
        times.setdefault(str(obj), []).append(order)
    return clicks, n, pos, times


def iou_and_simulated_clicks(pred, labels_full, inverse_map, labels_qv, coords_qv, current_num_clicks, n_obj=None):
    """One click round of eval_multi_obj.py:140-160 with ONE host read-back: the full-resolution IoU counts
    (mean_iou_scene) and the simulated next clicks (get_simulated_clicks, eval protocol) are both computed on the
    device and their two small records come back together, so the GPU never idles between them.
    -> ((mean IoU tensor, {object id: IoU}), (new_clicks, n, new_click_pos, new_click_time))."""
    if n_obj is None:
        n_obj = int(max(int(labels_full.max()), int(pred.max())) + 1) if labels_full.numel() else 1
    counts = ops.scene_iou_counts(pred.to(torch.int32).contiguous(), None if inverse_map is None else inverse_map.contiguous(),
                                  labels_full.to(torch.int32).contiguous(), n_obj)
    top_n = -1 if current_num_clicks == 0 else 1
    rec = ops.click_simulate(pred.to(torch.int32).contiguous(), labels_qv.to(torch.int32).contiguous(),
                             coords_qv.float().contiguous(), top_n=top_n, perm=None, max_new=64)
    both = torch.cat([counts.reshape(-1), rec.to(torch.int64)]).tolist()               # the one read-back of the round
    cnt, rec = both[:3 * n_obj], both[3 * n_obj:]
    ious = {}
    for o in range(1, n_obj):
        inter, npred, nlab = cnt[3 * o:3 * o + 3]
        if nlab:
            ious[o] = inter / (npred + nlab - inter)
    mean = sum(ious.values()) / len(ious) if ious else float("nan")
    n = rec[0]
    if n == 0:
        return (torch.tensor(mean), ious), (None, None, None, None)
    picks = [(rec[1 + 3 * i], rec[2 + 3 * i]) for i in range(n)]
    random.shuffle(picks)                                                              # utils/seg.py:127
    clicks, pos, times = {}, {}, {}
    for order, (row, obj) in enumerate(picks):
        clicks.setdefault(str(obj), []).append(int(row))
        pos.setdefault(str(obj), []).append(coords_qv[row])
        times.setdefault(str(obj), []).append(order)
    return (torch.tensor(mean), ious), (clicks, n, pos, times)


def extend_clicks(current_clicks, current_clicks_time, new_clicks, new_click_time):
    """utils/seg.py:229-239 (host dictionaries)."""
    current_click_num = sum(len(c) for c in current_clicks_time.values())
    for obj_id, click_ids in new_clicks.items():
        current_clicks[obj_id].extend(click_ids)
        current_clicks_time[obj_id].extend([t + current_click_num for t in new_click_time[obj_id]])
    return current_clicks, current_clicks_time
