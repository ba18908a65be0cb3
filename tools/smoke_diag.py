"""Per-layer deviation of the smoke scene's mask logits from the fp64 oracle, fp32 SIMT path vs tensor-core path."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import agile3d_b200  # noqa: E402
from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from oracle import me_ref  # noqa: E402
from oracle.agile3d_ref import build_ref_model  # noqa: E402

args = default_args()
model = agile3d_b200.build_model(args).eval()
sd = synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=11)
model.load_state_dict(sd)
model = model.cuda()
sc = make_scene(4000, 0.02, seed=42, n_box=8)
clicks, times, _ = make_clicks(sc, 2, 2, 1, seed=0)
coords = agile3d_b200.utils.batched_coordinates([sc["coords"]])
ref = build_ref_model(args).eval()
ref.load_state_dict(sd)
ref = ref.double()
with torch.no_grad():
    rx = me_ref.SparseTensor(coordinates=coords, features=torch.from_numpy(sc["feats"]).double())
    rh = ref.forward_backbone(rx, torch.from_numpy(sc["raw_coords"]).double())
    ro = ref.forward_mask(*rh, [clicks], [times])
rl = [a["pred_masks"][0].numpy() for a in ro["aux_outputs"]] + [ro["pred_masks"][0].numpy()]
rf = rh[0].F.numpy()
for algo, name in ((ops.ALGO_SIMT, "fp32 simt"), (ops.ALGO_AUTO, "tensor core")):
    model.backbone.algo = algo
    x = agile3d_b200.SparseTensor(coordinates=coords, features=torch.from_numpy(sc["feats"]), device="cuda")
    h = model.forward_backbone(x, raw_coordinates=torch.from_numpy(sc["raw_coords"]).cuda())
    out = model.forward_mask(*h, click_idx=[clicks], click_time_idx=[times])
    gl = [a["pred_masks"][0].cpu().numpy() for a in out["aux_outputs"]] + [out["pred_masks"][0].cpu().numpy()]
    fe = float(np.abs(h[0].F.cpu().numpy() - rf).max() / np.abs(rf).max())
    errs = [float(np.abs(g - r).max() / np.abs(r).max()) for g, r in zip(gl, rl)]
    flips = [int((g.argmax(1) != r.argmax(1)).sum()) for g, r in zip(gl, rl)]
    print(f"{name}: backbone features {fe:.2e}; logits per layer {['%.2e' % e for e in errs]}; label differences per layer {flips}")
