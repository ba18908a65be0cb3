"""Shared helpers for the test-suite (oracle construction, golden loading)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["g1500_k2", "g3000_k3", "g2500_k1_5cm"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["clicks"] = json.loads(str(g["clicks"]))
    g["times"] = json.loads(str(g["times"]))
    g["wseed"] = int(g["wseed"])
    return g


def layout():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


def oracle_model(wseed, dtype=torch.float32):
    from agile3d_b200.weights import default_args, synth_state_dict
    from oracle.agile3d_ref import build_ref_model

    m = build_ref_model(default_args()).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth_state_dict(shapes, seed=wseed))
    return m.to(dtype)


def oracle_forward(model, coords, feats, raw, clicks, times, dtype=torch.float32):
    """coords [N,4] int, feats [N,3], raw [N,3]; clicks/times lists of dicts -> (pcd, logits-per-layer-per-scene)."""
    from oracle import me_ref as ME

    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).to(dtype))
    with torch.no_grad():
        pcd, aux, co, pos = model.forward_backbone(x, torch.as_tensor(raw).to(dtype))
        out = model.forward_mask(pcd, aux, co, pos, clicks, times)
    per_layer = [a["pred_masks"] for a in out["aux_outputs"]] + [out["pred_masks"]]
    return pcd, aux, pos, per_layer


from oracle.compare import decision_forced_errors, rel_err  # noqa: E402,F401  (re-exported for the tests)


def oracle_train_step(model, coords, feats, raw, clicks, times, targets, dtype=torch.float32):
    """One train-mode step of the CPU oracle: batch-statistics BatchNorm, SetCriterion with click loss weights, the
    engine.py:128 weighted sum, backward.  targets: list of int arrays per scene.
    -> (loss dict, total, {param name: grad}, per-scene loss weights, output dict)."""
    from agile3d_b200.weights import default_args
    from oracle import criterion_ref as CR
    from oracle import me_ref as ME

    args = default_args()
    model.train()
    model.zero_grad()
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).to(dtype))
    rawt = torch.as_tensor(raw).to(dtype)
    pcd, aux, co, pos = model.forward_backbone(x, rawt)
    out = model.forward_mask(pcd, aux, co, pos, clicks, times)
    weights = []
    for b, r in enumerate(co[1]):
        ids = [int(i) for _, v in clicks[b].items() for i in v]
        weights.append(CR.click_loss_weights(rawt[torch.from_numpy(r)], ids))
    tg = [torch.as_tensor(t).long() for t in targets]
    loss_dict = CR.criterion(out, tg, weights)
    total = CR.total_loss(loss_dict, CR.weight_dict(args))
    total.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return loss_dict, total, grads, weights, out
