"""Pure-torch CPU restatement of the MinkowskiEngine v0.5.4 symbols AGILE3D touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: the real
MinkowskiEngine (NVIDIA/MinkowskiEngine git HEAD = v0.5.4, installed by the
reference's installation.md:30) is a third-party dependency that is not vendored
under /root/reference and cannot be installed here, so this file restates its
published semantics as written down in SURVEY.md Appendix A.  Each symbol cites the
reference call site that depends on it.

The module can be registered as ``MinkowskiEngine`` (``install_as_minkowski()``) so
that the *unmodified* reference model files run on top of it; that is how
tests/golden/make_golden.py produces the committed golden vectors.

Semantics restated (Appendix A):
  A.1  sparse_quantize: floor(coords / q), first-occurrence unique, inverse map
       (datasets/InterMultiObj3DSegDataset.py:67-71)
  A.2  batched_coordinates: prepend batch index (…Dataset.py:129)
  A.3  kernel offsets, x fastest; odd ks centred, even ks one-sided
       (models/modules/common.py:125-155)
  A.4  stride-2 conv: coarse set = unique(floor(c / 2ts) * 2ts), canonical row order =
       first occurrence while scanning fine rows (models/res16unet.py:51-59)
  A.5  transposed conv onto the cached finer map (models/modules/common.py:158-188)
  A.6  1x1 conv = dense matmul (models/resnet.py:109-116, models/agile3d.py:43-45)
  A.7  BatchNorm1d over all voxels, ReLU, +=, cat (models/modules/resnet_block.py:48-64)
  A.9  AvgPooling(2,2): mean over existing children (models/agile3d.py:71,172-173)
  A.10 SparseTensor ctor keeps row order of unique input (engine.py:47-51)
"""
from __future__ import annotations

import sys
import types
from enum import Enum

import numpy as np
import torch
import torch.nn as nn

# --------------------------------------------------------------------------- coords


def pack_keys(c: np.ndarray) -> np.ndarray:
    """(b, x, y, z) int rows -> one int64 key per row (16 bits each, xyz biased)."""
    c = np.asarray(c).astype(np.int64)
    if c.size and (np.abs(c[:, 1:]).max() >= 32768 or c[:, 0].min() < 0 or c[:, 0].max() >= 32768):
        raise ValueError("coordinate out of the +-32767 voxel range of the oracle key packing")
    return (c[:, 0] << 48) | ((c[:, 1] + 32768) << 32) | ((c[:, 2] + 32768) << 16) | (c[:, 3] + 32768)


def kernel_offsets(kernel_size: int, tensor_stride: int, dilation: int = 1) -> np.ndarray:
    """Appendix A.3: kernel index -> (dx, dy, dz), x fastest."""
    ks = int(kernel_size)
    offs = np.zeros((ks ** 3, 3), dtype=np.int64)
    for idx in range(ks ** 3):
        r = idx
        for axis in range(3):
            j = r % ks
            r //= ks
            if ks % 2 == 1:
                offs[idx, axis] = (j - ks // 2) * dilation * tensor_stride
            else:
                offs[idx, axis] = j * dilation * tensor_stride
    return offs


class CoordinateMapKey:
    def __init__(self, tensor_stride: int, tag: str = ""):
        self.tensor_stride = int(tensor_stride)
        self.tag = tag

    def get_tensor_stride(self):
        return [self.tensor_stride] * 3

    def _k(self):
        return (self.tensor_stride, self.tag)

    def __hash__(self):
        return hash(self._k())

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and self._k() == other._k()

    def __repr__(self):
        return f"CoordinateMapKey(stride={self.tensor_stride})"


class CoordinateManager:
    """Coordinate maps per tensor stride plus cached kernel maps (ME: CoordinateManager)."""

    def __init__(self):
        self.coords: dict[CoordinateMapKey, np.ndarray] = {}
        self._sorted: dict[CoordinateMapKey, tuple[np.ndarray, np.ndarray]] = {}
        self._kmaps: dict[tuple, np.ndarray] = {}
        self._parents: dict[tuple, np.ndarray] = {}

    # -- maps
    def insert(self, coords: np.ndarray, tensor_stride: int = 1) -> CoordinateMapKey:
        key = CoordinateMapKey(tensor_stride)
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        keys = pack_keys(coords)
        if np.unique(keys).shape[0] != keys.shape[0]:
            raise ValueError("SparseTensor coordinates must be unique (run sparse_quantize first)")
        self.coords[key] = coords
        return key

    def _lookup(self, key):
        if key not in self._sorted:
            k = pack_keys(self.coords[key])
            order = np.argsort(k, kind="stable")
            self._sorted[key] = (k[order], order)
        return self._sorted[key]

    def find_rows(self, key, query: np.ndarray) -> np.ndarray:
        """Row index in map `key` of each query coordinate, -1 where absent."""
        sk, order = self._lookup(key)
        q = pack_keys(query)
        pos = np.searchsorted(sk, q)
        pos_c = np.minimum(pos, sk.shape[0] - 1)
        hit = sk[pos_c] == q
        return np.where(hit, order[pos_c], -1).astype(np.int32)

    def stride(self, in_key: CoordinateMapKey, factor: int = 2) -> CoordinateMapKey:
        """Appendix A.4: coarse map, first-occurrence row order."""
        out_key = CoordinateMapKey(in_key.tensor_stride * factor)
        if out_key in self.coords:
            return out_key
        c = self.coords[in_key].astype(np.int64)
        new_ts = in_key.tensor_stride * factor
        coarse = c.copy()
        coarse[:, 1:] = np.floor_divide(c[:, 1:], new_ts) * new_ts
        k = pack_keys(coarse)
        _, first = np.unique(k, return_index=True)
        first = np.sort(first)
        self.coords[out_key] = coarse[first].astype(np.int32)
        return out_key

    def parent_rows(self, fine_key, coarse_key) -> np.ndarray:
        ck = (fine_key, coarse_key)
        if ck not in self._parents:
            c = self.coords[fine_key].astype(np.int64)
            ts = coarse_key.tensor_stride
            par = c.copy()
            par[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
            rows = self.find_rows(coarse_key, par)
            assert (rows >= 0).all()
            self._parents[ck] = rows
        return self._parents[ck]

    # -- kernel maps: nbr[k, o] = input row feeding output row o through offset k, or -1
    def kernel_map(self, in_key, out_key, kernel_size: int, dilation: int = 1) -> np.ndarray:
        ck = (in_key, out_key, int(kernel_size), int(dilation), "fwd")
        if ck not in self._kmaps:
            offs = kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
            out_c = self.coords[out_key].astype(np.int64)
            nbr = np.full((offs.shape[0], out_c.shape[0]), -1, dtype=np.int32)
            for k in range(offs.shape[0]):
                q = out_c.copy()
                q[:, 1:] += offs[k]
                nbr[k] = self.find_rows(in_key, q)
            self._kmaps[ck] = nbr
        return self._kmaps[ck]

    def kernel_map_transposed(self, in_key, out_key, kernel_size: int) -> np.ndarray:
        """Appendix A.5: the stride-2 map with in/out swapped; in = coarse, out = fine."""
        ck = (in_key, out_key, int(kernel_size), 1, "tr")
        if ck not in self._kmaps:
            assert kernel_size == 2, "only kernel 2 / stride 2 transposed convs are on the path"
            fine = self.coords[out_key].astype(np.int64)
            ts_f = out_key.tensor_stride
            par = self.parent_rows(out_key, in_key)
            coarse = self.coords[in_key].astype(np.int64)
            d = (fine[:, 1:] - coarse[par, 1:]) // ts_f          # each in {0,1}
            kidx = d[:, 0] + 2 * d[:, 1] + 4 * d[:, 2]
            nbr = np.full((8, fine.shape[0]), -1, dtype=np.int32)
            nbr[kidx, np.arange(fine.shape[0])] = par
            self._kmaps[ck] = nbr
        return self._kmaps[ck]


# --------------------------------------------------------------------------- tensor


class SparseTensor:
    """ME.SparseTensor: features [N,C] on a coordinate map (engine.py:47-51)."""

    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, device=None, **_):
        if device is not None:
            features = features.to(device)
        self.F = features
        if coordinates is not None:
            coords = coordinates.detach().cpu().numpy() if torch.is_tensor(coordinates) else np.asarray(coordinates)
            assert coords.ndim == 2 and coords.shape[1] == 4, "coordinates must be [N,4] (b,x,y,z)"
            assert coords.shape[0] == features.shape[0]
            self.coordinate_manager = CoordinateManager()
            self.coordinate_map_key = self.coordinate_manager.insert(coords, tensor_stride)
        else:
            assert coordinate_manager is not None and coordinate_map_key is not None
            self.coordinate_manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key
            assert self.coordinate_manager.coords[coordinate_map_key].shape[0] == features.shape[0]

    # ME attribute surface used by models/agile3d.py
    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.coords[self.coordinate_map_key]).to(self.F.device)

    @property
    def coordinates(self):
        return self.C

    @property
    def features(self):
        return self.F

    @property
    def D(self):
        return 3

    @property
    def device(self):
        return self.F.device

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def shape(self):
        return self.F.shape

    def _batch_rows(self):
        b = self.coordinate_manager.coords[self.coordinate_map_key][:, 0]
        nb = int(b.max()) + 1 if b.size else 0
        return [np.nonzero(b == i)[0] for i in range(nb)]

    @property
    def decomposed_features(self):
        return [self.F[torch.from_numpy(r)] for r in self._batch_rows()]

    @property
    def decomposed_coordinates(self):
        c = self.C
        return [c[torch.from_numpy(r)][:, 1:] for r in self._batch_rows()]

    def _like(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def __add__(self, other):
        assert other.coordinate_map_key == self.coordinate_map_key
        return self._like(self.F + other.F)

    def __iadd__(self, other):                      # models/modules/resnet_block.py:61
        assert other.coordinate_map_key == self.coordinate_map_key
        self.F = self.F + other.F
        return self

    def __repr__(self):
        return f"SparseTensor(N={self.F.shape[0]}, C={self.F.shape[1]}, {self.coordinate_map_key})"


def cat(*tensors):
    """MinkowskiOps.cat: channel concat on identical maps (models/res16unet.py:257)."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    k0 = tensors[0].coordinate_map_key
    assert all(t.coordinate_map_key == k0 for t in tensors)
    return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


# --------------------------------------------------------------------------- utils


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    """Appendix A.1."""
    is_np = isinstance(coordinates, np.ndarray)
    c = coordinates if is_np else coordinates.detach().cpu().numpy()
    if quantization_size is not None:
        disc = np.floor(c / quantization_size).astype(np.int32)
    else:
        disc = np.floor(c).astype(np.int32)
    k = pack_keys(np.concatenate([np.zeros((disc.shape[0], 1), np.int32), disc], 1))
    _, first, inv = np.unique(k, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-unique slot -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    unique_map = torch.from_numpy(first[order].astype(np.int64))
    inverse_map = torch.from_numpy(rank[inv.reshape(-1)].astype(np.int64))
    if return_maps_only:
        return (unique_map, inverse_map) if return_inverse else unique_map
    out_c = disc[unique_map.numpy()]
    if not is_np:
        out_c = torch.from_numpy(out_c)
    res = [out_c]
    if features is not None:
        res.append(features[unique_map.numpy() if isinstance(features, np.ndarray) else unique_map])
    if labels is not None:
        res.append(labels[unique_map.numpy() if isinstance(labels, np.ndarray) else unique_map])
    if return_index:
        res.append(unique_map)
    if return_inverse:
        res.append(inverse_map)
    return res[0] if len(res) == 1 else tuple(res)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """Appendix A.2."""
    rows = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c) if not torch.is_tensor(c) else c)
        c = torch.floor(c.double()).to(dtype) if c.is_floating_point() else c.to(dtype)
        rows.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c], dim=1))
    out = torch.cat(rows, 0) if rows else torch.zeros((0, 4), dtype=dtype)
    return out.to(device) if device is not None else out


# --------------------------------------------------------------------------- layers


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False,
                 region_type=RegionType.HYPER_CUBE, region_offsets=None, expand_coordinates=False,
                 axis_types=None, dimension=-1):
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        self.region_type, self.axis_types, self.dimension = region_type, axis_types, dimension


def _scalar(v):
    if isinstance(v, (list, tuple)):
        assert len(set(v[:3])) == 1, "anisotropic kernels are not on the AGILE3D path"
        return int(v[0])
    return int(v)


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, dimension=-1, transposed=False, **_):
        super().__init__()
        assert dimension == 3
        self.in_channels, self.out_channels = in_channels, out_channels
        self.ks, self.st, self.dil = _scalar(kernel_size), _scalar(stride), _scalar(dilation)
        self.transposed = transposed
        K = self.ks ** 3
        shape = (in_channels, out_channels) if (K == 1 and self.st == 1) else (K, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        # Appendix A.8 (ME reset_parameters)
        n = (out_channels if transposed else in_channels) * K
        s = 1.0 / (n ** 0.5)
        with torch.no_grad():
            self.kernel.uniform_(-s, s)
            if self.bias is not None:
                self.bias.uniform_(-s, s)

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        W = self.kernel
        if W.dim() == 2:                                                  # A.6
            out_key, out = in_key, x.F @ W
        else:
            if self.transposed:                                           # A.5
                assert self.st == 2 and in_key.tensor_stride % 2 == 0
                out_key = CoordinateMapKey(in_key.tensor_stride // 2)
                assert out_key in cm.coords, "transposed conv needs the cached finer map"
                nbr = cm.kernel_map_transposed(in_key, out_key, self.ks)
            else:
                out_key = cm.stride(in_key, self.st) if self.st > 1 else in_key   # A.4
                nbr = cm.kernel_map(in_key, out_key, self.ks, self.dil)           # A.3
            n_out = cm.coords[out_key].shape[0]
            out = x.F.new_zeros((n_out, self.out_channels))
            for k in range(nbr.shape[0]):
                sel = np.nonzero(nbr[k] >= 0)[0]
                if sel.size == 0:
                    continue
                o_idx = torch.from_numpy(sel)
                i_idx = torch.from_numpy(nbr[k][sel].astype(np.int64))
                out.index_add_(0, o_idx, x.F[i_idx] @ W[k])
        if self.bias is not None:
            out = out + self.bias
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    """models/modules/common.py:146-155."""

    def __init__(self, *a, **kw):
        super().__init__(*a, transposed=False, **kw)


class MinkowskiConvolutionTranspose(_ConvBase):
    """models/modules/common.py:179-188."""

    def __init__(self, *a, **kw):
        super().__init__(*a, transposed=True, **kw)


class MinkowskiBatchNorm(nn.Module):
    """Appendix A.7: nn.BatchNorm1d on .F, parameters under .bn.* (common.py:20-22)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return x._like(self.bn(x.F))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._like(torch.relu(x.F))


class MinkowskiAvgPooling(nn.Module):
    """Appendix A.9 (models/agile3d.py:71,173)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        self.ks, self.st = _scalar(kernel_size), _scalar(stride)
        assert self.ks == 2 and self.st == 2

    def forward(self, x):
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        out_key = cm.stride(in_key, 2)
        par = torch.from_numpy(cm.parent_rows(in_key, out_key).astype(np.int64))
        n_out = cm.coords[out_key].shape[0]
        s = x.F.new_zeros((n_out, x.F.shape[1])).index_add_(0, par, x.F)
        cnt = x.F.new_zeros((n_out, 1)).index_add_(0, par, x.F.new_ones((x.F.shape[0], 1)))
        return SparseTensor(s / cnt, coordinate_map_key=out_key, coordinate_manager=cm)


class _Unreached(nn.Module):
    """Referenced only inside never-called factory functions (common.py:24-28,231,252)."""

    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, *a, **kw):
        raise NotImplementedError(f"{type(self).__name__} is not on the AGILE3D hot path")


class MinkowskiInstanceNorm(_Unreached):
    pass


class MinkowskiSumPooling(_Unreached):
    pass


class MinkowskiAvgUnpooling(_Unreached):
    pass


# --------------------------------------------------------------------------- install


def install_as_minkowski():
    """Register this module as ``MinkowskiEngine`` (+ the three submodules the reference imports)."""
    me = types.ModuleType("MinkowskiEngine")
    me.__version__ = "0.5.4-oracle"
    for name in ("SparseTensor", "MinkowskiConvolution", "MinkowskiConvolutionTranspose", "MinkowskiBatchNorm",
                 "MinkowskiInstanceNorm", "MinkowskiReLU", "MinkowskiAvgPooling", "MinkowskiSumPooling",
                 "MinkowskiAvgUnpooling", "KernelGenerator", "RegionType", "MinkowskiNetwork",
                 "CoordinateManager", "CoordinateMapKey", "cat"):
        setattr(me, name, globals()[name])
    ops = types.ModuleType("MinkowskiEngine.MinkowskiOps")
    ops.cat, ops.SparseTensor = cat, SparseTensor
    pool = types.ModuleType("MinkowskiEngine.MinkowskiPooling")
    pool.MinkowskiAvgPooling = MinkowskiAvgPooling
    utils = types.ModuleType("MinkowskiEngine.utils")
    utils.sparse_quantize, utils.batched_coordinates = sparse_quantize, batched_coordinates
    me.MinkowskiOps, me.MinkowskiPooling, me.utils = ops, pool, utils
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiEngine.MinkowskiOps"] = ops
    sys.modules["MinkowskiEngine.MinkowskiPooling"] = pool
    sys.modules["MinkowskiEngine.utils"] = utils
    return me
