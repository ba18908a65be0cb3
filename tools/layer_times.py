"""Per-launch CUDA-event times of every library call of one eval step (ops.Profiler records in launch order), so the
spconv layers can be ranked at the bench batch size.  Usage: python tools/layer_times.py [--batch B] [--voxels N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxels", type=int, default=150000)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = agile3d_b200.build_model(default_args()).eval()
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev)
c, f, r, ck, tm = collate(make_inputs(a.batch, 2000, a.voxels))
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step():
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    return model.forward_mask(*model.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)


for _ in range(3):
    step()
torch.cuda.synchronize()
acc = None
REP = 3
for _ in range(REP):
    prof = ops.Profiler()
    ops.set_profiler(prof)
    step()
    ops.set_profiler(None)
    torch.cuda.synchronize()
    rows = [(n, b, fl, e0.elapsed_time(e1)) for n, b, fl, e0, e1 in prof.records]
    if acc is None:
        acc = [list(x) for x in rows]
    else:
        for x, y in zip(acc, rows):
            x[3] += y[3]
print(f"# batch {a.batch}, {c.shape[0]} voxels; mean of {REP} steps")
print(f"{'idx':>4} {'family':<10} {'ms':>8} {'MB':>9} {'GB/s':>8} {'GFLOP':>9} {'TFLOP/s':>8}")
tot = {}
for i, (n, b, fl, ms) in enumerate(acc):
    ms /= REP
    tot[n] = tot.get(n, 0.0) + ms
    if n in ("c2s", "s2c_mask") and i > 120:
        continue
    print(f"{i:4d} {n:<10} {ms:8.4f} {b / 1e6:9.2f} {b / ms / 1e6:8.1f} {fl / 1e9:9.2f} {fl / ms / 1e9:8.2f}")
print({k: round(v, 3) for k, v in tot.items()})
