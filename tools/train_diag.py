"""Per-layer deviation of the tensor-core (bf16x3) train-mode forward from the exact-fp32 one, on a golden scene:
max|z_tc - z_fp32| relative to max|z| and to the smallest per-channel std (what batch-statistics BatchNorm divides by)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import agile3d_b200  # noqa: E402
from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from helpers import load_golden  # noqa: E402


def run(algo, g):
    m = agile3d_b200.build_model(default_args())
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=g["wseed"]))
    m = m.cuda().train()
    m.backbone.algo = algo
    x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(g["coords"]), features=torch.as_tensor(g["feats"]), device="cuda")
    y, fmaps, maps, (tape, stem, _) = m.backbone.train_forward(x)
    recs = []
    for kind, r in tape:
        recs += [r] if kind == "conv" else [q for q in r if q is not None]
    return [(r["name"], r["z"].clone(), r["y"].clone()) for r in recs], y


g = load_golden("train_g1200_k2")
a, ya = run(ops.ALGO_SIMT, g)
b, yb = run(ops.ALGO_AUTO, g)
for (n, za, ya_), (_, zb, yb_) in zip(a, b):
    ez = float((za - zb).abs().max())
    std = za.std(0)
    mean = za.mean(0).abs()
    print(f"{n:28s} z rel {ez / float(za.abs().max()):.2e}  vs min-std {ez / float(std.min()):.2e}  "
          f"max |mean|/std {float((mean / std).max()):6.1f}  y rel {float((ya_ - yb_).abs().max()) / float(ya_.abs().max()):.2e}")
print("backbone out rel", float((ya - yb).abs().max()) / float(ya.abs().max()))
