"""Measured deviations of the GPU training step from the golden vector / the fp64 oracle (numbers behind the
tolerances in tests/test_gpu_train.py).  Usage: python tools/train_err.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, oracle_model, oracle_train_step, rel_err  # noqa: E402
import test_gpu_train as T  # noqa: E402

g = load_golden("train_g1200_k2")
names = json.loads(str(g["grad_names"]))
loss_names = json.loads(str(g["loss_names"]))
for algo in (1, 0):
    m = T._gpu_train_model(g["wseed"], algo)
    ld, total, grads, out = T._gpu_train_step(m, g["coords"], g["feats"], g["raw_coords"], g["clicks"], g["times"], [g["targets"]])
    got = np.array([float(ld[k].detach()) for k in loss_names])
    gn = np.array([float(grads[n].double().norm()) for n in names])
    print(f"golden algo={algo}: loss maxdiff {np.abs(got - g['loss_values']).max():.3e}  total-norm rel "
          f"{abs(np.sqrt((gn ** 2).sum()) - float(g['grad_total_norm'])) / float(g['grad_total_norm']):.3e}  "
          f"norms maxdiff/max {np.abs(gn - g['grad_norms']).max() / g['grad_norms'].max():.3e}  head-bias "
          f"{rel_err(grads['lin_squeeze_head.bias'].numpy(), g['grad_head_bias']):.3e}  bn0.w "
          f"{rel_err(grads['backbone.bn0.bn.weight'].numpy(), g['grad_bn0_weight']):.3e}  logits "
          f"{rel_err(out['pred_masks'][0].detach().cpu().numpy()[::4], g['logits_last']):.3e}")

from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
for n_vox in (1300, 6000):
    scs, clicks, times, targets = [], [], [], []
    for s in (dict(n=n_vox, seed=21, k=2, cpo=2, bg=1), dict(n=int(n_vox * 0.7), seed=22, k=1, cpo=3, bg=0)):
        sc = make_scene(s["n"], 0.02, seed=s["seed"], n_box=5)
        c, tm, lab = make_clicks(sc, s["k"], s["cpo"], s["bg"], seed=s["seed"])
        scs.append(sc); clicks.append(c); times.append(tm); targets.append(np.minimum(lab, len(c) - 1).astype(np.int32))
    coords = np.concatenate([np.concatenate([np.full((sc["coords"].shape[0], 1), b, np.int32), sc["coords"]], 1)
                             for b, sc in enumerate(scs)], 0)
    feats = np.concatenate([sc["feats"] for sc in scs], 0)
    raw = np.concatenate([sc["raw_coords"] for sc in scs], 0)
    rl, rtotal, rgrads, _, rout = oracle_train_step(oracle_model(7, torch.float64), coords, feats, raw, clicks, times, targets, torch.float64)
    r32 = oracle_train_step(oracle_model(7, torch.float32), coords, feats, raw, clicks, times, targets, torch.float32)
    cands = {"oracle-fp32": (r32[1], r32[2], r32[4])}
    for algo in (1, 0):
        m = T._gpu_train_model(7, algo)
        ld, total, grads, out = T._gpu_train_step(m, coords, feats, raw, clicks, times, targets)
        cands[f"gpu algo={algo}"] = (total, grads, out)
    for name, (total, grads, out) in cands.items():
        num = np.sqrt(sum(float((grads[n].double().cpu() - r).norm()) ** 2 for n, r in rgrads.items()))
        den = np.sqrt(sum(float(r.double().norm()) ** 2 for r in rgrads.values()))
        gmax = max(float(v.abs().max()) for v in rgrads.values())
        worst, wn = 0.0, ""
        for n, r in rgrads.items():
            e = float((grads[n].double().cpu() - r).abs().max()) / max(float(r.abs().max()), 1e-2 * gmax)
            if e > worst:
                worst, wn = e, n
        le = max(rel_err(out["pred_masks"][b].detach().cpu().numpy(), rout["pred_masks"][b].detach().numpy()) for b in range(2))
        print(f"batch2 N={coords.shape[0]} {name:14s}: total {float(total):.6f} vs {float(rtotal):.6f}  logits {le:.3e}  "
              f"grad L2 rel {num / den:.3e}  worst param {worst:.3e} ({wn})")
