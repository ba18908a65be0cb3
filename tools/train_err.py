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
import agile3d_b200  # noqa: E402
from oracle import me_ref as ME  # noqa: E402

# backbone only (no discrete decisions): features and gradients against the fp64 oracle
coords, feats, raw, clicks, times, targets = T._two_scenes()
R = torch.randn((coords.shape[0], 128), generator=torch.Generator().manual_seed(3), dtype=torch.float64)
for dt in (torch.float64, torch.float32):
    ref = oracle_model(7, dt).train()
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).to(dt))
    pcd_r, *_ = ref.forward_backbone(x, torch.as_tensor(raw).to(dt))
    (pcd_r.F * R.to(dt)).sum().backward()
    gr = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    if dt == torch.float64:
        rgrads, pcd64 = gr, pcd_r.F.detach()
    else:
        print("backbone-only oracle-fp32:", rel_err(pcd_r.F.detach().numpy(), pcd64.numpy()), T._grad_errors(gr, rgrads))
for algo in (1, 0):
    m = T._gpu_train_model(7, algo)
    xg = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device="cuda")
    pcd, *_ = m.forward_backbone(xg, torch.as_tensor(raw).cuda())
    (pcd.F * R.float().cuda()).sum().backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    print(f"backbone-only gpu algo={algo}:", rel_err(pcd.F.detach().cpu().numpy(), pcd64.numpy()), T._grad_errors(grads, rgrads))

# full-size scene: tensor-core (bf16x3) step against the exact-fp32 step, both on the GPU
sc = make_scene(150000, 0.02, seed=2000)
c, tm, lab = make_clicks(sc, 5, 2, 0, seed=2000)
coords = np.concatenate([np.zeros((sc["coords"].shape[0], 1), np.int32), sc["coords"]], 1)
res = {}
for algo in (1, 0):
    m = T._gpu_train_model(5, algo)
    ld, total, grads, out = T._gpu_train_step(m, coords, sc["feats"], sc["raw_coords"], [c], [tm], [lab.astype(np.int32)])
    res[algo] = (float(total), grads, out["pred_masks"][0].detach().cpu().numpy())
    del m
    torch.cuda.empty_cache()
num = np.sqrt(sum(float((res[0][1][n].double() - r.double()).norm()) ** 2 for n, r in res[1][1].items()))
den = np.sqrt(sum(float(r.double().norm()) ** 2 for r in res[1][1].values()))
print(f"150k voxels, tensor-core vs fp32 on the GPU: total {res[0][0]:.6f} vs {res[1][0]:.6f}  logits {rel_err(res[0][2], res[1][2]):.3e}  "
      f"grad L2 rel {num / den:.3e}")
