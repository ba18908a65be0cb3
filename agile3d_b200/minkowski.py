"""Host-side mirror of the MinkowskiEngine surface the reference's callers touch.

The reference callers (engine.py:47-51, eval_multi_obj.py:94-98, eval_single_obj.py:97-101,
interactive_tool/interactive_segmentation_user.py:191-195, datasets/*.py) only ever
  * voxelise with ``ME.utils.sparse_quantize`` and collate with ``ME.utils.batched_coordinates``,
  * build ``ME.SparseTensor(coordinates=[N,4] int32, features=[N,3] f32, device=...)``,
  * hand that tensor to ``model.forward_backbone``.
Everything else of MinkowskiEngine (coordinate manager, kernel maps, sparse convolution) happens
*inside* the model; here it lives in the CUDA library (csrc/) and is driven from ``backbone.py``.

``import agile3d_b200.minkowski as ME`` is therefore enough for the reference's callers.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["SparseTensor", "sparse_quantize", "batched_coordinates", "utils"]


def _voxel_keys(disc: np.ndarray) -> np.ndarray:
    """One sortable uint64 per integer voxel coordinate row (21 bits per axis, biased)."""
    d = disc.astype(np.int64) + (1 << 20)
    if d.size and (d.min() < 0 or d.max() >= (1 << 21)):
        raise ValueError("voxel coordinate outside +-2^20")
    return (d[:, 0].astype(np.uint64) << np.uint64(42)) | (d[:, 1].astype(np.uint64) << np.uint64(21)) \
        | d[:, 2].astype(np.uint64)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    """``ME.utils.sparse_quantize`` (datasets/InterMultiObj3DSegDataset.py:67-71).

    floor(coordinates / quantization_size) -> int32 voxels; one row per distinct voxel in order of
    first occurrence; ``unique_map`` = index of the first point of each voxel, ``inverse_map[i]`` =
    output row of point i.  Maps are torch int64; coordinates keep the input container type.
    """
    if torch.is_tensor(coordinates) and coordinates.is_cuda:
        # device path: floor-divide + first-occurrence hash unique + maps in csrc/click_ops.cu / coords.cu
        from . import ops
        vox, unique_map, inverse_map = ops.quantize_unique(coordinates, 1.0 if quantization_size is None else quantization_size)
        if return_maps_only:
            return (unique_map, inverse_map) if return_inverse else unique_map
        res = [vox[:, 1:].contiguous()]
        for extra in (features, labels):
            if extra is not None:
                res.append(extra[unique_map])
        if return_index:
            res.append(unique_map)
        if return_inverse:
            res.append(inverse_map)
        return res[0] if len(res) == 1 else tuple(res)
    is_np = isinstance(coordinates, np.ndarray)
    pts = coordinates if is_np else coordinates.detach().cpu().numpy()
    disc = np.floor(pts / quantization_size) if quantization_size is not None else np.floor(pts)
    disc = disc.astype(np.int32)
    keys = _voxel_keys(disc)
    order = np.argsort(keys, kind="stable")                    # stable: first point of a voxel leads its run
    sk = keys[order]
    head = np.ones(sk.shape[0], dtype=bool)
    head[1:] = sk[1:] != sk[:-1]
    first_pt = order[head]                                     # first-occurrence point per sorted voxel
    rank_of_sorted = np.argsort(np.argsort(first_pt, kind="stable"), kind="stable")
    run_id = np.cumsum(head) - 1
    inverse = np.empty(sk.shape[0], dtype=np.int64)
    inverse[order] = rank_of_sorted[run_id]
    unique = np.sort(first_pt).astype(np.int64)
    unique_map, inverse_map = torch.from_numpy(unique), torch.from_numpy(inverse)
    if return_maps_only:
        return (unique_map, inverse_map) if return_inverse else unique_map
    out_c = disc[unique]
    if not is_np:
        out_c = torch.from_numpy(out_c)
    res = [out_c]
    for extra in (features, labels):
        if extra is not None:
            res.append(extra[unique if isinstance(extra, np.ndarray) else unique_map])
    if return_index:
        res.append(unique_map)
    if return_inverse:
        res.append(inverse_map)
    return res[0] if len(res) == 1 else tuple(res)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """``ME.utils.batched_coordinates`` (datasets/InterMultiObj3DSegDataset.py:129): [sum Ni, 4] int32,
    column 0 = scene index, scenes concatenated in list order."""
    parts = []
    for b, c in enumerate(coords):
        c = torch.from_numpy(np.asarray(c)) if not torch.is_tensor(c) else c
        if c.is_floating_point():
            c = torch.floor(c)
        c = c.to(dtype)
        parts.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c], dim=1))
    out = torch.cat(parts, 0) if parts else torch.zeros((0, 4), dtype=dtype)
    return out.to(device) if device is not None else out


class SparseTensor:
    """``ME.SparseTensor(coordinates=, features=, device=)``: features [N,C] f32 on unique voxel
    coordinates [N,4] int32 (b,x,y,z), rows of a scene contiguous and in input order
    (SURVEY.md Appendix A.2/A.10).  The coordinate hash tables and kernel maps that
    MinkowskiEngine keeps in its CoordinateManager are built lazily on the GPU by
    ``agile3d_b200.backbone`` and cached on this object in ``.maps``.
    """

    def __init__(self, features, coordinates=None, device=None, tensor_stride=1, **kwargs):
        if coordinates is None:
            raise ValueError("agile3d_b200.SparseTensor needs coordinates=[N,4] (b,x,y,z)")
        if tensor_stride != 1:
            raise ValueError("input tensors are at tensor stride 1")
        coordinates = torch.as_tensor(coordinates)
        features = torch.as_tensor(features)
        if coordinates.dim() != 2 or coordinates.shape[1] != 4:
            raise ValueError("coordinates must be [N,4] (batch, x, y, z)")
        if coordinates.shape[0] != features.shape[0]:
            raise ValueError("coordinates and features disagree on N")
        dev = torch.device(device) if device is not None else features.device
        self._offsets = None
        if coordinates.device.type == "cpu" and coordinates.shape[0] > 0:
            self._offsets = self._offsets_from(coordinates[:, 0])      # free on the host, saves a device sync later
        self.C = coordinates.to(device=dev, dtype=torch.int32).contiguous()
        self.F = features.to(device=dev, dtype=torch.float32).contiguous()
        self.maps = None                      # filled by backbone.CoordinateMaps

    @staticmethod
    def _offsets_from(batch_col):
        b = batch_col.to(torch.int64)
        if b.numel() > 1 and bool((b[1:] < b[:-1]).any()):
            raise ValueError("rows of a scene must be contiguous and scenes in batch order (batched_coordinates)")
        counts = torch.bincount(b, minlength=int(b[-1]) + 1).tolist()
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        return offs

    def scene_offsets(self):
        """Row ranges of the scenes: offsets[b] .. offsets[b+1] (scenes are contiguous)."""
        if self._offsets is None:
            if self.C.is_cuda:
                # coordinates arrived on the GPU (eval_multi_obj.py:88-98): the row ranges come out of the same single
                # read-back as the level sizes of the coordinate maps (no 16 B/voxel copy to the host)
                from .backbone import CoordinateMaps
                self.maps = CoordinateMaps(self.C, want_offsets=True)
                self._offsets = self.maps.offsets
            else:
                self._offsets = self._offsets_from(self.C[:, 0])
        return self._offsets

    @property
    def coordinates(self):
        return self.C

    @property
    def features(self):
        return self.F

    @property
    def device(self):
        return self.F.device

    @property
    def D(self):
        return 3

    @property
    def shape(self):
        return self.F.shape

    def __repr__(self):
        return f"SparseTensor(N={self.F.shape[0]}, C={self.F.shape[1]}, device={self.F.device})"


class _Utils:
    sparse_quantize = staticmethod(sparse_quantize)
    batched_coordinates = staticmethod(batched_coordinates)


utils = _Utils()
