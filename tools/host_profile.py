"""cProfile of the host side of the eval forward at batch 1 (launch-bound): which python functions the ~3.9 ms per scene go to.
Usage: python tools/host_profile.py [--batch B]"""
import argparse
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = agile3d_b200.build_model(default_args()).eval()
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev)
c, f, r, ck, tm = collate(make_inputs(a.batch, 2000, 150000))
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step():
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    return model.forward_mask(*model.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumtime").print_stats(25)
