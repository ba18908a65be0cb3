"""Does a spatial row order let the dense-tile sparse convolution skip (tile, offset) stages?  
For a synthetic 150k-voxel scene: the 3x3x3 neighbour mask of every voxel, then for several row orders the number of
(128-row tile, offset) stages in which at least one row has a neighbour - the dense kernel's work - against the ideal
(pairs / 128).  Host numpy only.  Usage: python tools/order_analysis.py [voxels]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from agile3d_b200.scenes import make_scene  # noqa: E402


def part(v):
    v = v.astype(np.uint64) & np.uint64(0x1fffff)
    for sh, m in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
        v = (v | (v << np.uint64(sh))) & np.uint64(m)
    return v


def neighbour_mask(c):
    key = lambda a: (a[:, 0] + 4096) * (1 << 40) + (a[:, 1] + 4096) * (1 << 20) + (a[:, 2] + 4096)
    sk = np.sort(key(c))
    out = np.zeros((27, c.shape[0]), bool)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                q = key(c + np.array([dx, dy, dz]))
                pos = np.minimum(np.searchsorted(sk, q), len(sk) - 1)
                out[k] = sk[pos] == q
                k += 1
    return out


def stages(mask, perm, tile=128):
    m = mask[:, perm]
    nt = (m.shape[1] + tile - 1) // tile
    m = np.concatenate([m, np.zeros((27, nt * tile - m.shape[1]), bool)], 1)
    return int(m.reshape(27, nt, tile).any(2).sum()), nt


sc = make_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 150000, 0.02, seed=2000)
c = sc["coords"].astype(np.int64)
n = c.shape[0]
mask = neighbour_mask(c)
pairs = int(mask.sum())
print(f"voxels {n}, pairs {pairs}, mean neighbours per voxel {pairs / n:.2f} of 27, distinct 27-bit masks "
      f"{len(np.unique((mask * (1 << np.arange(27))[:, None]).sum(0)))}")
mk = part(c[:, 0]) | (part(c[:, 1]) << np.uint64(1)) | (part(c[:, 2]) << np.uint64(2))
bits = (mask.astype(np.uint64) * (np.uint64(1) << np.arange(27, dtype=np.uint64))[:, None]).sum(0)
orders = {"caller order (first occurrence)": np.arange(n), "Morton": np.argsort(mk, kind="stable"),
          "Morton blocks of 2^15 cells, rows sorted by mask inside": np.lexsort((mk, bits, mk >> np.uint64(15))),
          "whole scene sorted by mask (no locality)": np.argsort(bits, kind="stable")}
for name, perm in orders.items():
    s, nt = stages(mask, perm)
    print(f"{name:58s}: {s / nt:5.2f} of 27 stages per 128-row tile non-empty, MMA rows useful {pairs / (s * 128):.2f}")
print(f"pair-packed (csrc/spconv_pk.cu): groups of 128 pairs per 256-row super tile: ~30 per super tile = 15 per 128 rows "
      f"(mean fill 0.67)")
