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


def _by_path(name, path):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


TRAIN_CASES = [
    # one scene: the reference's CPU branch (agile3d.py:146-150,197-202) is only correct for batch size 1
    dict(name="train_g1200_k2", scenes=[dict(n=1200, voxel=0.02, seed=11, n_box=5, k=2, cpo=2, bg=1)], wseed=4),
]


def main_train():
    """Train-mode step of the UNMODIFIED reference (models/*.py, models/criterion.py, utils/seg.py) on oracle.me_ref:
    forward_backbone + forward_mask in train mode (batch-statistics BatchNorm), SetCriterion with
    cal_click_loss_weights, the engine.py:128 weighted sum, backward.  Stores the loss dictionary, the total
    gradient norm and per-parameter gradient norms / sums."""
    crit_mod = _by_path("ref_criterion", "/root/reference/models/criterion.py")
    seg_mod = _by_path("ref_seg", "/root/reference/utils/seg.py")
    args = default_args()
    criterion = crit_mod.build_mask_criterion(args)
    for case in TRAIN_CASES:
        torch.manual_seed(0)
        model = models.build_model(args)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(synth_state_dict(shapes, seed=case["wseed"]))
        model.train()
        scs, clicks, times, labels = [], [], [], []
        for s in case["scenes"]:
            sc = make_scene(s["n"], s["voxel"], seed=s["seed"], n_box=s["n_box"])
            c, t, l = make_clicks(sc, s["k"], s["cpo"], s["bg"], seed=s["seed"])
            scs.append(sc); clicks.append(c); times.append(t); labels.append(l)
        coords = ME.utils.batched_coordinates([sc["coords"] for sc in scs])
        feats = torch.from_numpy(np.concatenate([sc["feats"] for sc in scs], 0))
        raw = torch.from_numpy(np.concatenate([sc["raw_coords"] for sc in scs], 0))
        x = ME.SparseTensor(coordinates=coords, features=feats)
        pcd, aux, coordinates, pos = model.forward_backbone(x, raw_coordinates=raw)
        out = model.forward_mask(pcd, aux, coordinates, pos, click_idx=clicks, click_time_idx=times)
        targets = [torch.from_numpy(np.minimum(l, len(c) - 1)).float() for l, c in zip(labels, clicks)]
        weights = seg_mod.cal_click_loss_weights(coords[:, 0], raw, torch.cat(targets), clicks)
        loss_dict = criterion(out, targets, weights)
        total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
        total.backward()
        names = [n for n, p in model.named_parameters() if p.grad is not None]
        gn = np.array([float(dict(model.named_parameters())[n].grad.double().norm()) for n in names])
        gs = np.array([float(dict(model.named_parameters())[n].grad.double().sum()) for n in names])
        np.savez_compressed(
            os.path.join(ROOT, f"tests/golden/{case['name']}.npz"),
            coords=coords.numpy(), feats=feats.numpy(), raw_coords=raw.numpy(),
            targets=np.concatenate([t.numpy() for t in targets]).astype(np.int32),
            clicks=json.dumps(clicks), times=json.dumps(times), wseed=case["wseed"],
            loss_names=json.dumps(sorted(loss_dict)), loss_values=np.array([float(loss_dict[k]) for k in sorted(loss_dict)]),
            total=float(total), weights=np.concatenate([w.numpy() for w in weights]),
            grad_names=json.dumps(names), grad_norms=gn, grad_sums=gs,
            grad_total_norm=float(np.sqrt((gn ** 2).sum())),
            grad_head_bias=model.lin_squeeze_head.bias.grad.numpy(),
            grad_bn0_weight=model.backbone.bn0.bn.weight.grad.numpy(),
            logits_last=out["pred_masks"][0].detach().numpy()[::4].copy(),
            bn0_running_mean=model.backbone.bn0.bn.running_mean.numpy(),
        )
        print(case["name"], "N =", coords.shape[0], {k: round(float(v), 5) for k, v in loss_dict.items()},
              "total", float(total), "|g| =", float(np.sqrt((gn ** 2).sum())), "params with grad", len(names))


if __name__ == "__main__":
    if "--train-only" not in sys.argv:
        main()
    main_train()
