"""Runs the batched end-to-end test input through the eval backbone with a synchronise after every sparse conv and
prints the arguments of the first call that raises."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import agile3d_b200  # noqa: E402
from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.scenes import make_scene  # noqa: E402
import test_gpu_parity as T  # noqa: E402

sa = make_scene(6000, 0.02, seed=21, n_box=8)
sb = make_scene(9000, 0.05, seed=22, n_box=10)
coords = np.concatenate([np.concatenate([np.zeros((sa["coords"].shape[0], 1), np.int32), sa["coords"]], 1),
                         np.concatenate([np.ones((sb["coords"].shape[0], 1), np.int32), sb["coords"]], 1)], 0)
feats = np.concatenate([sa["feats"], sb["feats"]], 0)
orig = ops.spconv_fwd
count = [0]


def traced(x, nbr, weight, out, *a, **kw):
    count[0] += 1
    info = (f"#{count[0]} x {tuple(x.shape)} stride {x.stride()} off {x.storage_offset()} ptr%128={x.data_ptr() % 128} "
            f"nbr {None if nbr is None else tuple(nbr.shape)} w {tuple(weight.shape)} out {tuple(out.shape)} stride {out.stride()} "
            f"kw { {k: (v if not torch.is_tensor(v) else tuple(v.shape)) for k, v in kw.items() if k != 'weight_tc'} }")
    if nbr is not None:
        info += f" nbr max {int(nbr.max())} min {int(nbr.min())}"
    try:
        r = orig(x, nbr, weight, out, *a, **kw)
        torch.cuda.synchronize()
        return r
    except Exception as e:  # noqa: BLE001
        print("FAILED", info, str(e)[:100], flush=True)
        raise


ops.spconv_fwd = traced
import agile3d_b200.backbone as bb  # noqa: E402
bb.ops.spconv_fwd = traced
m = T._gpu_model(7)
x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device="cuda")
y, fmaps, maps = m.backbone(x)
torch.cuda.synchronize()
print("backbone ok,", count[0], "convs; level sizes", maps.sizes)
