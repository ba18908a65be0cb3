"""One training step (forward, criterion, backward, clip + AdamW) between cudaProfilerStart/Stop, for ncu
(`--profile-from-start off`).  Usage: python tools/profile_train_step.py [--voxels N] [--batch B]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200.optim import FlatAdamW  # noqa: E402
from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxels", type=int, default=150000)
ap.add_argument("--batch", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
margs = default_args()
model = agile3d_b200.build_model(margs)
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev).train()
criterion = agile3d_b200.build_criterion(margs)
opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=0.1)
batch, targets = [], []
for i in range(a.batch):
    sc = make_scene(a.voxels, 0.02, seed=2000 + i)
    ck, tm, lab = make_clicks(sc, 5, 2, 0, seed=2000 + i)
    batch.append((sc, ck, tm))
    targets.append(torch.from_numpy(lab.astype(np.int32)).to(dev))
c, f, r, ck, tm = collate(batch)
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step():
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    out = model.forward_mask(*model.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)
    w = agile3d_b200.cal_click_loss_weights(c[:, 0], r, None, ck)
    ld = criterion(out, targets, w)
    total = sum(ld[k] * criterion.weight_dict[k] for k in ld if k in criterion.weight_dict)
    opt.zero_grad()
    total.backward()
    opt.step()
    return total


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one training step:", c.shape[0], "voxels")
