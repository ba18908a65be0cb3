"""Decoder parameter gradients of one tensor-core-mode training step with the X^T dY contractions on tcgen05 vs on the
fp32 SIMT kernel (everything else identical), and the 6-step loss history of tests::test_training_reduces_the_loss in
both settings."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import agile3d_b200  # noqa: E402
from agile3d_b200 import model as M, ops  # noqa: E402
from agile3d_b200.optim import FlatAdamW  # noqa: E402
from agile3d_b200.weights import default_args  # noqa: E402
from helpers import load_golden  # noqa: E402
import test_gpu_train as T  # noqa: E402

g = load_golden("train_g1200_k2")
orig = M._xt_dy


def simt_xt_dy(xs, dy):
    out = None
    for x in xs:
        out = ops.spconv_bwd_weight(x, None, dy, 1, dweight=out, accumulate=out is not None)
    return out


res = {}
for mode in ("tc", "simt"):
    M._xt_dy = orig if mode == "tc" else simt_xt_dy
    m = T._gpu_train_model(g["wseed"])
    ld, total, grads, out = T._gpu_train_step(m, g["coords"], g["feats"], g["raw_coords"], g["clicks"], g["times"], [g["targets"]])
    res[mode] = {n: v.double().cpu() for n, v in grads.items() if not n.startswith("backbone.")}
    print(mode, "total loss", float(total))
worst = sorted(((float((res["tc"][n] - res["simt"][n]).norm() / max(float(res["simt"][n].norm()), 1e-30)), n) for n in res["tc"]), reverse=True)
print("decoder-side gradients, tc vs simt contraction: worst relative L2 differences")
for e, n in worst[:8]:
    print(f"   {e:.3e} {n}")

for mode in ("tc", "simt"):
    M._xt_dy = orig if mode == "tc" else simt_xt_dy
    m = T._gpu_train_model(g["wseed"])
    criterion = agile3d_b200.build_criterion(default_args())
    opt = FlatAdamW(m.parameters(), lr=2e-4, weight_decay=1e-4, max_norm=0.1)
    x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(g["coords"]), features=torch.as_tensor(g["feats"]), device="cuda")
    raw = torch.as_tensor(g["raw_coords"]).cuda()
    tg = [torch.as_tensor(g["targets"]).cuda()]
    weights = agile3d_b200.cal_click_loss_weights(x.C[:, 0], raw, tg[0], g["clicks"])
    hist = []
    for _ in range(8):
        opt.zero_grad()
        o = m.forward_mask(*m.forward_backbone(x, raw), g["clicks"], g["times"])
        l = criterion(o, tg, weights)
        total = sum(l[k] * criterion.weight_dict[k] for k in l)
        total.backward()
        opt.step()
        hist.append(round(float(total), 3))
    print(mode, "loss history", hist)
