"""Known-answer tests for the MinkowskiEngine restatement (oracle/me_ref.py) — SURVEY.md §8(c) list.
These pin the *stated* semantics (Appendix A) since the reference ships no tests of its own."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import me_ref as ME


def _st(coords3, feats, batch=0):
    c = np.concatenate([np.full((len(coords3), 1), batch), np.asarray(coords3)], 1).astype(np.int32)
    return ME.SparseTensor(coordinates=torch.from_numpy(c), features=torch.as_tensor(feats, dtype=torch.float32))


def test_kernel_offset_order_x_fastest():
    o = ME.kernel_offsets(3, 1)
    assert o[0].tolist() == [-1, -1, -1] and o[1].tolist() == [0, -1, -1]
    assert o[13].tolist() == [0, 0, 0] and o[26].tolist() == [1, 1, 1]
    o2 = ME.kernel_offsets(2, 4)
    assert o2[0].tolist() == [0, 0, 0] and o2[1].tolist() == [4, 0, 0] and o2[7].tolist() == [4, 4, 4]
    o5 = ME.kernel_offsets(5, 1)
    assert o5[0].tolist() == [-2, -2, -2] and o5[62].tolist() == [0, 0, 0]


def test_conv3_one_hot_on_five_voxels():
    # voxels: centre + its +x, -x, +y, +z neighbours; feature = one-hot voxel id
    pts = [(5, 5, 5), (6, 5, 5), (4, 5, 5), (5, 6, 5), (5, 5, 6)]
    x = _st(pts, np.eye(5))
    conv = ME.MinkowskiConvolution(5, 27, kernel_size=3, dimension=3)
    with torch.no_grad():
        conv.kernel.zero_()
        for k in range(27):
            conv.kernel[k, :, k] = 1.0          # output channel k counts "input voxel seen through offset k"
    y = conv(x).F.detach().numpy()
    # output row 0 (centre) sees itself via k=13, +x via k=14, -x via k=12, +y via k=16, +z via k=22
    assert sorted(np.nonzero(y[0])[0].tolist()) == [12, 13, 14, 16, 22]
    # row 1 (+x voxel) sees the centre through (-1,0,0)=k12, the +y voxel through (-1,+1,0)=k15, +z through (-1,0,+1)=k21
    assert sorted(np.nonzero(y[1])[0].tolist()) == [12, 13, 15, 21]


def test_dense_equivalence_with_conv3d():
    g = 6
    zz, yy, xx = np.meshgrid(np.arange(g), np.arange(g), np.arange(g), indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1)
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((pts.shape[0], 4)).astype(np.float32)
    conv = ME.MinkowskiConvolution(4, 5, kernel_size=3, dimension=3)
    y = conv(_st(pts, feats)).F.detach()
    dense = torch.zeros(1, 4, g, g, g)
    dense[0, :, pts[:, 2], pts[:, 1], pts[:, 0]] = torch.from_numpy(feats).T
    # kernel index k = ix + 3*iy + 9*iz (x fastest) -> conv3d weight [Cout, Cin, kz, ky, kx]; cross-correlation
    w = conv.kernel.detach().reshape(3, 3, 3, 4, 5).permute(4, 3, 0, 1, 2)
    ref = F.conv3d(dense, w, padding=1)[0]
    got = torch.zeros_like(ref)
    got[:, pts[:, 2], pts[:, 1], pts[:, 0]] = y.T
    assert torch.allclose(got, ref, atol=1e-4)


def test_stride2_down_then_transposed_up():
    pts = [(0, 0, 0), (1, 0, 0), (1, 1, 1), (2, 3, 0), (5, 4, 2), (4, 4, 2)]
    x = _st(pts, np.arange(6, dtype=np.float32)[:, None] + 1)
    down = ME.MinkowskiConvolution(1, 1, kernel_size=2, stride=2, dimension=3)
    up = ME.MinkowskiConvolutionTranspose(1, 1, kernel_size=2, stride=2, dimension=3)
    with torch.no_grad():
        down.kernel.fill_(1.0)
        up.kernel.copy_(torch.arange(8, dtype=torch.float32).reshape(8, 1, 1) + 1)   # W[k] = k+1
    d = down(x); d.F = d.F.detach()
    # canonical coarse order = first occurrence: (0,0,0) <- rows 0,1,2 ; (2,2,0) <- row 3 ; (4,4,2) <- rows 4,5
    assert d.C[:, 1:].tolist() == [[0, 0, 0], [2, 2, 0], [4, 4, 2]]
    assert d.F[:, 0].tolist() == [1 + 2 + 3, 4, 5 + 6]
    u = up(d); u.F = u.F.detach()
    assert u.coordinate_map_key == x.coordinate_map_key
    # fine voxel f gets parent feature * W[k(f - parent)]
    k_of = [0, 1, 7, 2, 1, 0]
    parent = [6, 6, 6, 4, 11, 11]
    assert u.F[:, 0].tolist() == [p * (k + 1) for p, k in zip(parent, k_of)]


def test_negative_coordinates_floor_division():
    x = _st([(-1, -1, -1), (-2, 0, 1), (0, 0, 0)], np.ones((3, 1)))
    d = ME.MinkowskiConvolution(1, 1, kernel_size=2, stride=2, dimension=3)(x)
    assert d.C[:, 1:].tolist() == [[-2, -2, -2], [-2, 0, 0], [0, 0, 0]]


def test_avg_pool_means_existing_children():
    x = _st([(0, 0, 0), (1, 1, 0), (2, 0, 0)], [[2.0], [4.0], [10.0]])
    p = ME.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(x)
    assert p.F[:, 0].tolist() == [3.0, 10.0]


def test_sparse_quantize_first_occurrence_and_inverse():
    pts = np.array([[0.031, 0.0, 0.0], [0.001, 0.0, 0.0], [0.039, 0.001, 0.0], [0.0, 0.021, 0.0]], np.float32)
    c, um, im = ME.sparse_quantize(pts, quantization_size=0.02, return_index=True, return_inverse=True)
    assert c.tolist() == [[1, 0, 0], [0, 0, 0], [0, 1, 0]]
    assert um.tolist() == [0, 1, 3] and im.tolist() == [0, 1, 0, 2]
    assert um.dtype == torch.int64 and isinstance(c, np.ndarray) and c.dtype == np.int32
    # the package's own host implementation follows the same rule
    from agile3d_b200.minkowski import sparse_quantize, batched_coordinates
    rng = np.random.default_rng(1)
    big = rng.random((5000, 3)).astype(np.float32) * 0.5
    a = ME.sparse_quantize(big, quantization_size=0.02, return_index=True, return_inverse=True)
    b = sparse_quantize(big, quantization_size=0.02, return_index=True, return_inverse=True)
    assert np.array_equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.equal(ME.batched_coordinates([a[0], a[0][:7]]), batched_coordinates([b[0], b[0][:7]]))
    assert np.array_equal(big[a[1].numpy()][a[2].numpy()] // 0.02 * 0 + a[0][a[2].numpy()], np.floor(big / 0.02).astype(np.int32))


def test_duplicate_coordinates_rejected():
    import pytest
    with pytest.raises(ValueError):
        _st([(0, 0, 0), (0, 0, 0)], np.ones((2, 1)))
