"""Generate tests/golden/*.npz by running the UNMODIFIED reference model files.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It registers oracle.me_ref under the name ``MinkowskiEngine`` (the real one is not installable here,
see oracle/me_ref.py), imports the reference's own ``models`` package from /root/reference, builds
``models.build_model(args)`` with the reference's default flags (main.py:34-58), loads the deterministic
synthetic checkpoint ``agile3d_b200.weights.synth_state_dict(seed)`` and records, per case, the inputs
and the outputs of ``forward_backbone`` + ``forward_mask`` (eval mode, fp32, CPU).

What the vectors pin: the backbone graph (res16unet.py), forward_backbone/forward_mask/mask_module
(agile3d.py), the fourier/time encodings (position_embedding.py) and the decoder layers
(attention_block.py) exactly as the reference wrote them.  What they do not pin: MinkowskiEngine's
kernels themselves (restated, parity unpinned).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import me_ref  # noqa: E402

me_ref.install_as_minkowski()
sys.path.insert(0, "/root/reference")
import models  # noqa: E402  (the reference's package)
import MinkowskiEngine as ME  # noqa: E402  (= oracle.me_ref)

from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402

CASES = [
    # name, target voxels, voxel size, scene seed, n_box, objects, clicks/object, bg clicks, weight seed
    dict(name="g1500_k2", n=1500, voxel=0.02, seed=3, n_box=6, k=2, cpo=2, bg=1, wseed=1),
    dict(name="g3000_k3", n=3000, voxel=0.02, seed=5, n_box=8, k=3, cpo=3, bg=0, wseed=2),
    dict(name="g2500_k1_5cm", n=2500, voxel=0.05, seed=9, n_box=6, k=1, cpo=3, bg=2, wseed=3),
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    args = default_args()
    model = models.build_model(args).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(ROOT, "tests/golden/state_dict_layout.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f, indent=0, sort_keys=True)
    for case in CASES:
        model.load_state_dict(synth_state_dict(shapes, seed=case["wseed"]))
        sc = make_scene(case["n"], case["voxel"], seed=case["seed"], n_box=case["n_box"])
        clicks, times, labels = make_clicks(sc, case["k"], case["cpo"], case["bg"], seed=case["seed"])
        coords = ME.utils.batched_coordinates([sc["coords"]])
        x = ME.SparseTensor(coordinates=coords, features=torch.from_numpy(sc["feats"]))
        with torch.no_grad():
            pcd, aux, coordinates, pos = model.forward_backbone(x, raw_coordinates=torch.from_numpy(sc["raw_coords"]))
            out = model.forward_mask(pcd, aux, coordinates, pos, click_idx=[clicks], click_time_idx=[times])
        layers = [a["pred_masks"][0] for a in out["aux_outputs"]] + [out["pred_masks"][0]]
        np.savez_compressed(
            os.path.join(ROOT, f"tests/golden/{case['name']}.npz"),
            coords=coords.numpy(), feats=sc["feats"], raw_coords=sc["raw_coords"], labels=labels,
            clicks=json.dumps(clicks), times=json.dumps(times), wseed=case["wseed"],
            pcd_features=pcd.F.numpy()[::4].copy(),            # every 4th row keeps the fixture small
            pos_enc=pos[4][0][0].numpy()[::16].copy(),
            level_sizes=np.array([a.F.shape[0] for a in aux]),
            aux_feat_sums=np.array([float(a.F.double().sum()) for a in aux]),
            logits=np.stack([l.numpy() for l in layers], 0),
        )
        print(case["name"], "N =", coords.shape[0], "levels", [a.F.shape[0] for a in aux],
              "labels", layers[-1].argmax(1).bincount().tolist())


if __name__ == "__main__":
    main()
