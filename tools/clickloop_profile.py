"""cProfile of the host side of the C5 click loop (one scene, 192 rounds): where the ~1.5 ms of host time per round go.
Usage: python tools/clickloop_profile.py"""
import copy
import cProfile
import os
import pstats
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200 import interactive, ops  # noqa: E402
from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
model = agile3d_b200.build_model(default_args()).eval()
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev)
K, max_per_obj = 10, 20
sc = make_scene(80000, 0.05, seed=5000, outdoor=True)
_, _, lab = make_clicks(sc, K, 1, 0, seed=5000)
lab = torch.from_numpy(np.minimum(lab, K).astype(np.int64))
c = agile3d_b200.utils.batched_coordinates([sc["coords"]]).to(dev)
f, r = torch.from_numpy(sc["feats"]).to(dev), torch.from_numpy(sc["raw_coords"]).to(dev)
lab_full, inv, lab = lab[sc["inverse_map"]].to(dev), sc["inverse_map"].to(dev), lab.to(dev)


def protocol():
    random.seed(0)
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    h = model.forward_backbone(x, raw_coordinates=r)
    click_idx = {str(o): [] for o in range(K + 1)}
    click_time = copy.deepcopy(click_idx)
    n_clicks, nv = 0, c.shape[0]
    while n_clicks <= K * max_per_obj:
        if n_clicks == 0:
            pred = ops.click_pred(None, nv, K + 1, torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
        else:
            out = model.forward_mask(*h, click_idx=[click_idx], click_time_idx=[click_time])
            rows = torch.tensor([v for o in range(K + 1) for v in click_idx[str(o)]], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
            objs = torch.tensor([o for o in range(K + 1) for _ in click_idx[str(o)]], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
            pred = ops.click_pred(out["pred_masks"][0], nv, K + 1, rows, objs)
        (_, _), (new, _, _, new_t) = interactive.iou_and_simulated_clicks(pred, lab_full, inv, lab, r, n_clicks, n_obj=K + 1)
        if new is not None:
            click_idx, click_time = interactive.extend_clicks(click_idx, click_time, new, new_t)
        n_clicks += K if n_clicks == 0 else 1


protocol()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
protocol()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
