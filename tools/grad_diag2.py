"""Block-by-block gradients of the fp32 GPU backbone backward against the fp64 oracle: for every BasicBlock, the
gradient arriving at its output (dout) and the gradient it passes to its input (dx)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_model  # noqa: E402
import test_gpu_train as T  # noqa: E402
import agile3d_b200  # noqa: E402
from agile3d_b200.backbone import Res16UNet34C  # noqa: E402
from oracle import me_ref as ME  # noqa: E402
from oracle.agile3d_ref import RefBasicBlock  # noqa: E402

coords, feats, raw, *_ = T._two_scenes()
R = torch.randn((coords.shape[0], 128), generator=torch.Generator().manual_seed(3), dtype=torch.float64)
ref = oracle_model(7, torch.float64).train()
cap = {}


def hook(name):
    def f(mod, inp, out):
        inp[0].F.retain_grad()
        out.F.retain_grad()
        cap[name] = (inp[0].F, out.F)
    return f


for n, mod in ref.named_modules():
    if isinstance(mod, RefBasicBlock):
        mod.register_forward_hook(hook(n.replace("backbone.", "")))
x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).double())
pcd_r, *_ = ref.forward_backbone(x, torch.as_tensor(raw).double())
(pcd_r.F * R).sum().backward()

order = []
for stage in (8, 7, 6, 5, 4, 3, 2, 1):
    nb = len(getattr(ref.backbone, f"block{stage}"))
    order += [f"block{stage}.{b}" for b in reversed(range(nb))]

rec = []
orig = Res16UNet34C._block_train_bwd


extra = []


def patched(self, blk_rec, dout, W, grads):
    d_in = dout.detach().clone()
    r1, r2, rd = blk_rec
    y_before = r2["y"].detach().clone()
    z_before = r2["z"].detach().clone()
    dbeta_torch = (d_in * (r2["y"] > 0)).sum(0).double().cpu()
    dx = orig(self, blk_rec, dout, W, grads)
    extra.append(dict(y=y_before.double().cpu(), dbeta_torch=dbeta_torch, dbeta_kernel=grads[r2["bn"] + ".bn.bias"].double().cpu(),
                      bn=r2["bn"], y_changed=float((r2["y"] - y_before).abs().max()), z_changed=float((r2["z"] - z_before).abs().max()),
                      x_ptr=r1["x"].data_ptr(), x_shape=tuple(r1["x"].shape), x_stride=tuple(r1["x"].stride())))
    rec.append((d_in.double().cpu(), dx.detach().clone().double().cpu()))
    return dx


Res16UNet34C._block_train_bwd = patched
m = T._gpu_train_model(7, 1)
xg = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device="cuda")
pcd, *_ = m.forward_backbone(xg, torch.as_tensor(raw).cuda())
(pcd.F * R.float().cuda()).sum().backward()


def rel(a, b):
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


print(f"{'block':<10} {'dout err':>10} {'dx err':>10}   shapes")
for name, (d_in, dx) in zip(order, rec):
    xin, yout = cap[name]
    print(f"{name:<10} {rel(d_in, yout.grad):10.2e} {rel(dx, xin.grad):10.2e}   {tuple(d_in.shape)} -> {tuple(dx.shape)}")

rg = {n: p.grad.double() for n, p in ref.named_parameters() if p.grad is not None}
for name, e in list(zip(order, extra))[:4]:
    yo = cap[name][1].detach()
    print(name, "y vs oracle", rel(e["y"], yo), " dbeta torch-vs-oracle", rel(e["dbeta_torch"], rg["backbone." + e["bn"] + ".bn.bias"]),
          " kernel-vs-oracle", rel(e["dbeta_kernel"], rg["backbone." + e["bn"] + ".bn.bias"]), " y/z changed during block bwd", e["y_changed"], e["z_changed"],
          " x", e["x_shape"], e["x_stride"])

for name, e in list(zip(order, extra))[:3]:
    yo = cap[name][1].detach()
    yg = e["y"]
    mm = (yg > 0) != (yo > 0)
    idx = mm.nonzero()
    print(name, "mask mismatches", int(mm.sum()), "of", yo.numel(), " 0<y_or<1e-4:", int(((yo > 0) & (yo < 1e-4)).sum()),
          " 0<y_gpu<1e-4:", int(((yg > 0) & (yg < 1e-4)).sum()))
    for r, c in idx[:8].tolist():
        print("    row", r, "ch", c, "y_gpu", float(yg[r, c]), "y_or", float(yo[r, c]))
    if len(idx):
        rows = idx[:, 0]
        print("    mismatching rows: min", int(rows.min()), "max", int(rows.max()), "distinct", len(set(rows.tolist())),
              " channels distinct", len(set(idx[:, 1].tolist())))
