"""Per-parameter gradient errors of the fp32 GPU backbone backward against the fp64 oracle (and the fp32 oracle as the
yardstick) on the two tiny scenes of tests/test_gpu_train.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_model, rel_err  # noqa: E402
import test_gpu_train as T  # noqa: E402
import agile3d_b200  # noqa: E402
from oracle import me_ref as ME  # noqa: E402

coords, feats, raw, *_ = T._two_scenes()
R = torch.randn((coords.shape[0], 128), generator=torch.Generator().manual_seed(3), dtype=torch.float64)
res = {}
for dt in (torch.float64, torch.float32):
    ref = oracle_model(7, dt).train()
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).to(dt))
    pcd_r, *_ = ref.forward_backbone(x, torch.as_tensor(raw).to(dt))
    (pcd_r.F * R.to(dt)).sum().backward()
    res[dt] = ({n: p.grad.double() for n, p in ref.named_parameters() if p.grad is not None}, pcd_r.F.detach().double())
for algo in (1, 0):
    m = T._gpu_train_model(7, algo)
    xg = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device="cuda")
    pcd, *_ = m.forward_backbone(xg, torch.as_tensor(raw).cuda())
    (pcd.F * R.float().cuda()).sum().backward()
    res[algo] = ({n: p.grad.double().cpu() for n, p in m.named_parameters() if p.grad is not None}, pcd.F.detach().double().cpu())
g64, f64 = res[torch.float64]
order = list(g64)
for key in (torch.float32, 1, 0):
    g, f = res[key]
    print(f"== {key}: feature err {rel_err(f.numpy(), f64.numpy()):.3e}  all-grads {T._grad_errors(g, g64)}")
    rows = [(float((g[n] - g64[n]).norm() / max(float(g64[n].norm()), 1e-30)), i, n) for i, n in enumerate(order)]
    print("   first 12 in execution order:", " ".join(f"{e:.1e}" for e, _, _ in rows[:12]))
    for e, i, n in sorted(rows, reverse=True)[:8]:
        print(f"   {e:.3e}  #{i:3d} {n}  |g| {float(g64[n].norm()):.3e}")
    if key == 1:
        print("   all parameters, forward order (backward runs bottom-up):")
        for e, i, n in rows:
            print(f"     {e:.2e} {n}")
