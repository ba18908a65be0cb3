"""Deterministic synthetic scenes of the BASELINE.json shapes.

Host-side numpy only: this is the data generator for tests and bench.py, not part of
the hot path.  A scene is an axis-aligned room shell (floor + 4 walls) with `n_box`
axis-aligned boxes (5 visible faces each); points are sampled uniformly on the faces,
jittered, shifted to the origin and voxelised with the first-occurrence rule of
``sparse_quantize`` (reference: datasets/InterMultiObj3DSegDataset.py:50-75).
The room is rescaled until the voxel count is within +-2 % of the target.
"""
from __future__ import annotations

import numpy as np

from .minkowski import sparse_quantize


def _sample_rect(rng, origin, u, v, density):
    """Uniform points on the rectangle origin + a*u + b*v, a,b in [0,1]."""
    area = np.linalg.norm(u) * np.linalg.norm(v)
    n = max(4, int(area * density))
    ab = rng.random((n, 2))
    return origin[None, :] + ab[:, :1] * u[None, :] + ab[:, 1:] * v[None, :]


def _raw_scene(rng_seed, scale, voxel, n_box, outdoor):
    rng = np.random.default_rng(rng_seed)
    density = 1.5 / (voxel * voxel)                        # >= 1.5 points per voxel-face area
    if outdoor:
        L, W, H = 60.0 * scale, 60.0 * scale, 3.0
    else:
        L, W, H = 8.0 * scale, 6.0 * scale, 3.0 * min(scale, 1.0)
    ex, ey, ez = np.eye(3)
    pts, lab = [], []
    # room shell: label 0
    shell = [_sample_rect(rng, np.zeros(3), ex * L, ey * W, density)]
    if not outdoor:
        shell += [
            _sample_rect(rng, np.zeros(3), ex * L, ez * H, density),
            _sample_rect(rng, ey * W, ex * L, ez * H, density),
            _sample_rect(rng, np.zeros(3), ey * W, ez * H, density),
            _sample_rect(rng, ex * L, ey * W, ez * H, density),
        ]
    for s in shell:
        pts.append(s)
        lab.append(np.zeros(s.shape[0], np.int32))
    # boxes: label 1..n_box, 5 visible faces (no bottom)
    boxes = []
    for j in range(n_box):
        sz = rng.uniform(0.3, 1.5, 3) * (scale if not outdoor else 1.0)
        sz = np.maximum(sz, 4.0 * voxel)
        sz[2] = min(sz[2], H * 0.8)
        lo = np.array([rng.uniform(0.05 * L, 0.95 * L - sz[0]), rng.uniform(0.05 * W, 0.95 * W - sz[1]), 0.0])
        bx, by, bz = ex * sz[0], ey * sz[1], ez * sz[2]
        faces = [
            _sample_rect(rng, lo + bz, bx, by, density),           # top
            _sample_rect(rng, lo, bx, bz, density),
            _sample_rect(rng, lo + by, bx, bz, density),
            _sample_rect(rng, lo, by, bz, density),
            _sample_rect(rng, lo + bx, by, bz, density),
        ]
        for f in faces:
            pts.append(f)
            lab.append(np.full(f.shape[0], j + 1, np.int32))
        boxes.append((lo, sz))
    pts = np.concatenate(pts, 0)
    lab = np.concatenate(lab, 0)
    pts = pts + rng.normal(0.0, 0.2 * voxel, pts.shape)
    pts = (pts - pts.min(0, keepdims=True)).astype(np.float32)
    rgb = rng.random((pts.shape[0], 3), dtype=np.float32)
    return pts, rgb, lab, boxes


def make_scene(target_voxels: int, voxel_size: float = 0.02, seed: int = 0, n_box: int = 30,
               outdoor: bool = False, tol: float = 0.02):
    """Returns dict(coords int32[N,3], raw_coords f32[N,3], feats f32[N,3], labels int32[N], ...)."""
    # surface-area estimate of the starting scale, then fixed-point refinement N ~ scale^2
    base_area = (60.0 * 60.0) if outdoor else (8 * 6 + 2 * (8 + 6) * 3.0)
    scale = float(np.sqrt(max(target_voxels, 64) * voxel_size ** 2 / base_area))
    best = None
    for _ in range(12):
        pts, rgb, lab, boxes = _raw_scene(seed, scale, voxel_size, n_box, outdoor)
        coords, umap, imap = sparse_quantize(pts, quantization_size=voxel_size, return_index=True,
                                             return_inverse=True)
        n = coords.shape[0]
        best = (pts, rgb, lab, boxes, coords, umap, imap, scale)
        if abs(n - target_voxels) <= tol * target_voxels:
            break
        scale *= float(np.sqrt(target_voxels / n))
    pts, rgb, lab, boxes, coords, umap, imap, scale = best
    umap_np = umap.numpy()
    return {
        "coords": np.ascontiguousarray(coords, dtype=np.int32),
        "raw_coords": np.ascontiguousarray(pts[umap_np], dtype=np.float32),
        "feats": np.ascontiguousarray(rgb[umap_np], dtype=np.float32),
        "labels": np.ascontiguousarray(lab[umap_np], dtype=np.int32),
        "labels_full": lab,
        "inverse_map": imap,
        "voxel_size": voxel_size,
        "scale": scale,
        "seed": seed,
    }


def make_clicks(scene, num_obj: int, clicks_per_obj: int, num_bg_clicks: int = 0, seed: int = 0):
    """Click dicts in the reference convention: key '0' background, '1'..'K' objects.

    Object j's first click is the voxel nearest the centroid of the box-top voxels, further clicks are
    random voxels of the object; click time index = order of generation.  The scene labels are
    remapped so that the K chosen boxes become ids 1..K and everything else is background.
    """
    rng = np.random.default_rng(seed + 7919)
    lab = scene["labels"]
    raw = scene["raw_coords"]
    present = [j for j in np.unique(lab) if j != 0 and (lab == j).sum() >= clicks_per_obj + 1]
    assert len(present) >= num_obj, "scene has too few objects"
    chosen = present[:num_obj]
    new_lab = np.zeros_like(lab)
    clicks = {"0": []}
    times = {"0": []}
    t = 0
    for new_id, j in enumerate(chosen, start=1):
        rows = np.nonzero(lab == j)[0]
        new_lab[rows] = new_id
        top = rows[raw[rows, 2] >= raw[rows, 2].max() - 2 * scene["voxel_size"]]
        centre = raw[top].mean(0)
        first = int(top[np.argmin(((raw[top] - centre) ** 2).sum(1))])
        rest = [int(r) for r in rng.permutation(rows) if int(r) != first][: clicks_per_obj - 1]
        clicks[str(new_id)] = [first] + rest
        times[str(new_id)] = list(range(t, t + clicks_per_obj))
        t += clicks_per_obj
    bg_rows = np.nonzero(new_lab == 0)[0]
    for r in rng.permutation(bg_rows)[:num_bg_clicks]:
        clicks["0"].append(int(r))
        times["0"].append(t)
        t += 1
    return clicks, times, new_lab
