"""cProfile of the host side of the training step (4 x 150k voxels, launch-bound).  Usage: python tools/host_profile_train.py"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200.optim import FlatAdamW  # noqa: E402
from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate  # noqa: E402

dev = torch.device("cuda", 0)
margs = default_args()
model = agile3d_b200.build_model(margs)
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev).train()
criterion = agile3d_b200.build_criterion(margs)
opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=0.1)
batch, targets = [], []
for i in range(4):
    sc = make_scene(150000, 0.02, seed=2000 + i)
    clicks, times, lab = make_clicks(sc, 5, 2, 0, seed=2000 + i)
    batch.append((sc, clicks, times))
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


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(34)
st.sort_stats("cumtime").print_stats(30)
