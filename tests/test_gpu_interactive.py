"""GPU parity tests (-m gpu) of the interactive-loop kernels and the voxelisation front end (SURVEY.md 8(f) rows 1-2;
csrc/click_ops.cu, agile3d_b200/interactive.py) against the CPU oracle (oracle/seg_ref.py, itself pinned to the
unmodified utils/seg.py in tests/test_seg_ref_cpu.py) and against the host sparse_quantize."""
import random

import numpy as np
import pytest
import torch

from oracle import seg_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agile3d_b200._lib import lib
    lib()


def _scene(seed, n, k):
    from agile3d_b200.scenes import make_clicks, make_scene
    sc = make_scene(n, 0.05 if n < 100000 else 0.02, seed=seed, n_box=max(k + 2, 8))
    _, _, lab = make_clicks(sc, k, 1, 0, seed=seed)
    g = torch.Generator().manual_seed(seed)
    labels = torch.from_numpy(lab.astype(np.int64))
    pred = labels.clone()
    flip = torch.rand(labels.shape[0], generator=g) < 0.25
    pred[flip] = torch.randint(0, k + 1, (int(flip.sum()),), generator=g)
    return sc, torch.from_numpy(sc["raw_coords"]), labels, pred


@pytest.mark.parametrize("n,k", [(3000, 3), (20000, 6), (80000, 10)])
@pytest.mark.parametrize("mode", ["first", "later", "train"])
def test_simulated_clicks_vs_oracle(n, k, mode):
    from agile3d_b200 import interactive
    sc, xyz, labels, pred = _scene(n + k, n, k)
    if mode == "first":
        pred = torch.zeros_like(labels)                              # eval_multi_obj.py:120-121
    kw = dict(current_num_clicks=0 if mode == "first" else 5, training=mode == "train")
    random.seed(3)
    want = seg_ref.get_simulated_clicks(pred, labels, xyz, **kw)
    random.seed(3)
    got = interactive.get_simulated_clicks(pred.to(DEV), labels.to(DEV), xyz.to(DEV), **kw)
    assert got[1] == want[1]
    if got[0] != want[0]:
        # torch.cdist evaluates |a|^2 + |b|^2 - 2ab in fp32, the kernel (a-b)^2: two voxels of a cluster whose border distances
        # agree to rounding may swap.  Then the sizes of the picked voxels must agree to 1e-5.
        cl = seg_ref.error_clusters(pred, labels, xyz)
        for obj, rows in got[0].items():
            for r, r_ref in zip(rows, want[0][obj]):
                if r != r_ref:
                    inside = (pred != labels) & (labels == labels[r]) & (pred == pred[r])
                    d = torch.cdist(xyz[~inside], xyz[[r, r_ref]]).min(0)[0]
                    assert abs(float(d[0] - d[1])) < 1e-5 * float(d.max()), (obj, r, r_ref, d)
    else:
        assert got[3] == want[3]
        assert all(torch.equal(a.cpu(), b) for o in got[2] for a, b in zip(got[2][o], want[2][o]))


@pytest.mark.parametrize("n_clicks", [0, 7])
def test_fused_round_equals_separate_calls(n_clicks):
    """iou_and_simulated_clicks (one read-back per round) == mean_iou_scene + get_simulated_clicks"""
    from agile3d_b200 import interactive
    sc, xyz, labels, pred = _scene(77, 20000, 6)
    inv = sc["inverse_map"]
    lab_full = labels[inv]
    p, l, lf, iv, x = pred.to(DEV), labels.to(DEV), lab_full.to(DEV), inv.to(DEV), xyz.to(DEV)
    random.seed(3)
    iou_a, ious_a = interactive.mean_iou_scene(p, lf, iv)
    a = interactive.get_simulated_clicks(p, l, x, n_clicks, training=False)
    random.seed(3)
    (iou_b, ious_b), b = interactive.iou_and_simulated_clicks(p, lf, iv, l, x, n_clicks)
    assert float(iou_a) == float(iou_b) and ious_a == ious_b
    assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3]
    assert all(torch.equal(u, v) for k in a[2] for u, v in zip(a[2][k], b[2][k]))


def test_nothing_wrong_returns_nones():
    from agile3d_b200 import interactive
    _, xyz, labels, _ = _scene(1, 3000, 3)
    assert interactive.get_simulated_clicks(labels.to(DEV), labels.to(DEV), xyz.to(DEV), 3, False) == (None, None, None, None)


def test_click_pred_and_iou_vs_oracle():
    from agile3d_b200 import interactive, ops
    sc, xyz, labels, _ = _scene(9, 20000, 5)
    g = torch.Generator().manual_seed(2)
    logits = torch.randn((labels.shape[0], 6), generator=g)
    logits[::7, 2] = logits[::7, 4] = 9.0                                                   # ties: first maximum wins
    rows = torch.tensor([5, 17, 400, 9999], dtype=torch.int32)
    objs = torch.tensor([1, 0, 5, 3], dtype=torch.int32)
    want = logits.argmax(1)
    want[rows.long()] = objs.long()
    got = ops.click_pred(logits.to(DEV), labels.shape[0], 6, rows.to(DEV), objs.to(DEV))
    assert torch.equal(got.cpu().long(), want)
    assert int(ops.click_pred(None, 1000, 3, rows[:2].to(DEV), objs[:2].to(DEV)).sum()) == 1   # round 0: zeros + clicks
    inv = sc["inverse_map"]
    lab_full = torch.from_numpy(sc["labels_full"].astype(np.int64)).clamp(max=5)
    m_ref, d_ref = seg_ref.mean_iou_scene(want[inv], lab_full)
    m, d = interactive.mean_iou_scene(got, lab_full.to(DEV), inv.to(DEV))
    assert abs(float(m) - m_ref) < 1e-6 and set(d) == set(d_ref) and all(abs(d[k] - d_ref[k]) < 1e-9 for k in d)


@pytest.mark.parametrize("n,voxel", [(5000, 0.05), (400000, 0.02)])
def test_sparse_quantize_on_device_equals_host(n, voxel):
    import agile3d_b200
    g = torch.Generator().manual_seed(n)
    pts = (torch.rand((n, 3), generator=g) * torch.tensor([6.0, 4.0, 2.5]) - 1.0).float()       # negative coordinates too
    feats = torch.rand((n, 3), generator=g)
    c_ref, f_ref, u_ref, i_ref = agile3d_b200.utils.sparse_quantize(pts.numpy(), features=feats.numpy(), return_index=True,
                                                                    return_inverse=True, quantization_size=voxel)
    c, f, u, i = agile3d_b200.utils.sparse_quantize(pts.to(DEV), features=feats.to(DEV), return_index=True, return_inverse=True,
                                                    quantization_size=voxel)
    assert torch.equal(c.cpu(), torch.from_numpy(c_ref)) and torch.equal(u.cpu(), u_ref) and torch.equal(i.cpu(), i_ref)
    assert torch.equal(f.cpu(), torch.from_numpy(f_ref))
    assert torch.equal(c[i.to(DEV)].cpu(), torch.floor(pts / np.float32(voxel)).int())       # round trip through inverse_map
