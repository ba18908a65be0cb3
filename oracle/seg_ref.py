"""TEST INFRASTRUCTURE (oracle): CPU restatement of the interactive-loop helpers of the reference, utils/seg.py.

    mean_iou_scene            utils/seg.py:9-17,44-59
    error_clusters / ranked_clicks / get_simulated_clicks     utils/seg.py:94-226
    extend_clicks             utils/seg.py:229-239

Pinned against the UNMODIFIED reference functions (imported by file path) in tests/test_seg_ref_cpu.py.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import this module; the product path is csrc/click_ops.cu.
"""
import random

import torch


def mean_iou_scene(pred, labels):
    """utils/seg.py:44-59: mean over the objects present in labels (id != 0) of |p & l| / |p | l|."""
    ids = [int(i) for i in torch.unique(labels) if int(i) != 0]
    out = {}
    for i in ids:
        p, l = pred == i, labels == i
        inter = int((p & l).sum())
        out[i] = inter / (int(p.sum()) + int(l.sum()) - inter)
    return (sum(out.values()) / len(out) if out else float("nan")), out


def error_clusters(pred, labels, xyz):
    """utils/seg.py:173-205: {cluster id 96 gt + 11 pred: (size, first voxel attaining it)} where the size of a cluster is
    the largest distance of one of its voxels to the nearest voxel outside the cluster (measure_error_size, :154-171)."""
    pred, labels = pred.float(), labels.float()
    wrong = pred != labels
    cid = labels * 96 + pred * 11
    out = {}
    for c in torch.unique(cid[wrong]).tolist():
        inside = wrong & (cid == c)
        if int(inside.sum()) == 0 or int((~inside).sum()) == 0:
            continue
        d = torch.cdist(xyz[~inside], xyz[inside]).min(dim=0)[0]          # per cluster voxel: distance to the border
        rows = torch.nonzero(inside).squeeze(1)
        out[int(c)] = (float(d.max()), int(rows[int(torch.nonzero(d == d.max())[0, 0])]))
    return out


def ranked_clicks(pred, labels, xyz, current_num_clicks=None, training=True):
    """clusters by size, descending (python's stable sort over ascending cluster ids, utils/seg.py:207), cut as the
    reference does (:209-218) -> [(voxel row, object id = ground truth of that voxel)]"""
    cl = error_clusters(pred, labels, xyz)
    order = sorted(cl, key=lambda c: cl[c][0], reverse=True)
    if training:
        order = order[:int((torch.unique(labels) != 0).sum())]
    elif current_num_clicks != 0:
        order = order[:1]
    return [(cl[c][1], int(labels[cl[c][1]])) for c in order]


def get_simulated_clicks(pred, labels, xyz, current_num_clicks=None, training=True):
    """utils/seg.py:173-226 including the random.shuffle of the selected clusters (:127)."""
    picks = ranked_clicks(pred, labels, xyz, current_num_clicks, training)
    if not picks:
        return None, None, None, None
    random.shuffle(picks)
    clicks, pos, times = {}, {}, {}
    for order, (row, obj) in enumerate(picks):
        clicks.setdefault(str(obj), []).append(row)
        pos.setdefault(str(obj), []).append(xyz[row])
        times.setdefault(str(obj), []).append(order)
    return clicks, len(picks), pos, times


def extend_clicks(current_clicks, current_clicks_time, new_clicks, new_click_time):
    n = sum(len(c) for c in current_clicks_time.values())
    for obj, ids in new_clicks.items():
        current_clicks[obj].extend(ids)
        current_clicks_time[obj].extend([t + n for t in new_click_time[obj]])
    return current_clicks, current_clicks_time
