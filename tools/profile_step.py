"""One eval forward (forward_backbone + forward_mask) of the headline scene between cudaProfilerStart/Stop, for
ncu (`--profile-from-start off`).  Usage: python tools/profile_step.py [--voxels N] [--batch B] [--algo 0|1|2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxels", type=int, default=150000)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--algo", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = agile3d_b200.build_model(default_args()).eval()
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev)
model.backbone.algo = a.algo
c, f, r, ck, tm = collate(make_inputs(a.batch, 2000, a.voxels))
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step():
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    return model.forward_mask(*model.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step:", c.shape[0], "voxels")
