"""The CPU oracle's TRAIN step (oracle/agile3d_ref.py + oracle/criterion_ref.py, torch autograd) against the golden
vector produced by the UNMODIFIED reference files (models/*.py, models/criterion.py, utils/seg.py) in train mode
(tests/golden/make_golden.py::main_train): loss dictionary, click loss weights, per-parameter gradient norms."""
import json

import numpy as np
import torch

from helpers import load_golden, oracle_model, oracle_train_step, rel_err


def _load():
    g = load_golden("train_g1200_k2")
    g["loss_names"] = json.loads(str(g["loss_names"]))
    g["grad_names"] = json.loads(str(g["grad_names"]))
    return g


def test_oracle_train_step_reproduces_reference():
    g = _load()
    m = oracle_model(g["wseed"])
    loss_dict, total, grads, weights, out = oracle_train_step(
        m, g["coords"], g["feats"], g["raw_coords"], g["clicks"], g["times"], [g["targets"]])
    assert sorted(loss_dict) == g["loss_names"]
    got = np.array([float(loss_dict[k]) for k in g["loss_names"]])
    assert np.abs(got - g["loss_values"]).max() < 1e-5
    assert abs(float(total) - float(g["total"])) < 1e-4
    assert rel_err(weights[0].numpy(), g["weights"]) < 1e-6
    assert rel_err(out["pred_masks"][0].detach().numpy()[::4], g["logits_last"]) < 1e-5
    assert sorted(grads) == sorted(g["grad_names"])
    gn = np.array([float(grads[n].double().norm()) for n in g["grad_names"]])
    # same arithmetic, same order of operations up to the restated glue -> gradient norms agree closely
    assert np.abs(gn - g["grad_norms"]).max() / g["grad_norms"].max() < 1e-3
    assert rel_err(grads["lin_squeeze_head.bias"].numpy(), g["grad_head_bias"]) < 1e-3
    assert rel_err(grads["backbone.bn0.bn.weight"].numpy(), g["grad_bn0_weight"]) < 2e-3
    assert rel_err(m.backbone.bn0.bn.running_mean.numpy(), g["bn0_running_mean"]) < 1e-5


def test_criterion_restatement_properties():
    """Known answers for the per-voxel loss: uniform logits, perfect prediction, weights scale linearly."""
    from oracle import criterion_ref as CR
    n, C = 50, 4
    t = torch.arange(n) % C
    w = torch.ones(n)
    uni = CR.criterion({"pred_masks": [torch.zeros(n, C)]}, [t], [w])
    assert abs(float(uni["loss_bce"]) - np.log(C)) < 1e-6
    # p_t = 1/C: dice = 1 - (2/C^2 + eps) / (2/C + eps)
    assert abs(float(uni["loss_dice"]) - (1 - (2 / C ** 2 + 1e-6) / (2 / C + 1e-6))) < 1e-6
    sharp = torch.nn.functional.one_hot(t, C).float() * 50
    perfect = CR.criterion({"pred_masks": [sharp]}, [t], [w])
    assert float(perfect["loss_bce"]) < 1e-6 and float(perfect["loss_dice"]) < 1e-6
    twice = CR.criterion({"pred_masks": [torch.zeros(n, C)]}, [t], [2 * w])
    assert abs(float(twice["loss_bce"]) - 2 * float(uni["loss_bce"])) < 1e-6
    xyz = torch.tensor([[0.0, 0, 0], [0.15, 0, 0], [1.0, 0, 0]])
    wts = CR.click_loss_weights(xyz, [0])
    assert torch.allclose(wts, torch.tensor([2.0, 1.4, 0.8]), atol=1e-6)
