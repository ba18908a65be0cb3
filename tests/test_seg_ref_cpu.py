"""Pins oracle/seg_ref.py (the checker of the device click simulation) against the UNMODIFIED reference utils/seg.py,
imported by file path.  Runs where /root/reference exists (this container); the GPU box only sees the oracle."""
import importlib.util
import os
import random

import numpy as np
import pytest
import torch

from oracle import seg_ref

REF = "/root/reference/utils/seg.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


def _ref():
    spec = importlib.util.spec_from_file_location("ref_seg", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _scene(seed, n=1500, k=4):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand((n, 3), generator=g) * torch.tensor([4.0, 3.0, 1.0])
    labels = (xyz[:, 0] * k / 4.0).long().clamp(0, k)               # slabs along x: objects 0..k
    pred = labels.clone()
    flip = torch.rand(n, generator=g) < 0.3
    pred[flip] = torch.randint(0, k + 1, (int(flip.sum()),), generator=g)
    return xyz, labels, pred


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("mode", ["first", "later", "train"])
def test_simulated_clicks_equal_reference(seed, mode):
    ref = _ref()
    xyz, labels, pred = _scene(seed)
    kw = dict(current_num_clicks=0 if mode == "first" else 3, training=mode == "train")
    random.seed(seed)
    a = ref.get_simulated_clicks(pred.clone(), labels.clone(), xyz, **kw)
    random.seed(seed)
    b = seg_ref.get_simulated_clicks(pred, labels, xyz, **kw)
    assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3]
    for k_ in a[2]:
        assert all(torch.equal(u, v) for u, v in zip(a[2][k_], b[2][k_]))


def test_first_round_all_zero_prediction():
    ref = _ref()
    xyz, labels, _ = _scene(5)
    pred = torch.zeros_like(labels)
    random.seed(1)
    a = ref.get_simulated_clicks(pred.clone(), labels.clone(), xyz, current_num_clicks=0, training=False)
    random.seed(1)
    b = seg_ref.get_simulated_clicks(pred, labels, xyz, current_num_clicks=0, training=False)
    assert a[0] == b[0] and a[1] == b[1] == int((torch.unique(labels) != 0).sum()) and a[3] == b[3]


def test_nothing_wrong_and_iou_and_extend():
    ref = _ref()
    xyz, labels, pred = _scene(7)
    assert seg_ref.get_simulated_clicks(labels, labels, xyz, 2, False) == (None, None, None, None)
    m_ref, d_ref = ref.mean_iou_scene(pred, labels)
    m, d = seg_ref.mean_iou_scene(pred, labels)
    assert abs(float(m_ref) - m) < 1e-6 and all(abs(d_ref[k] - d[k]) < 1e-6 for k in d_ref) and set(d) == set(d_ref)
    cur = {"0": [], "1": [5], "2": [7, 9]}
    tim = {"0": [], "1": [0], "2": [1, 2]}
    new, nt = {"2": [11], "0": [3]}, {"2": [0], "0": [1]}
    import copy
    a = ref.extend_clicks(copy.deepcopy(cur), copy.deepcopy(tim), new, nt)
    b = seg_ref.extend_clicks(copy.deepcopy(cur), copy.deepcopy(tim), new, nt)
    assert a == b
    from agile3d_b200 import interactive
    assert interactive.extend_clicks(copy.deepcopy(cur), copy.deepcopy(tim), new, nt) == a
