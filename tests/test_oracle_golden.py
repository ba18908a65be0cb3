"""The CPU oracle (oracle/agile3d_ref.py on oracle/me_ref.py) against the golden vectors produced by the
UNMODIFIED reference model files (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, layout, load_golden, oracle_forward, oracle_model, rel_err


def test_state_dict_layout_matches_reference():
    from agile3d_b200.weights import default_args
    from oracle.agile3d_ref import build_ref_model

    ours = {k: tuple(v.shape) for k, v in build_ref_model(default_args()).state_dict().items()}
    ref = layout()
    assert set(ours) == set(ref)
    assert all(ours[k] == ref[k] for k in ref)
    assert sum(int(np.prod(v)) for k, v in ref.items() if not k.endswith("num_batches_tracked")
               and "running" not in k and k != "pos_enc.gauss_B") == 39289760


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_outputs(name):
    g = load_golden(name)
    m = oracle_model(g["wseed"])
    pcd, aux, pos, per_layer = oracle_forward(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    assert [a.F.shape[0] for a in aux] == g["level_sizes"].tolist()
    # same arithmetic in the same order -> expect (near) bit equality with the reference run
    assert rel_err(pcd.F.numpy()[::4], g["pcd_features"]) < 1e-6
    assert rel_err(pos[0].numpy()[::16], g["pos_enc"]) < 1e-6
    for l in range(3):
        assert rel_err(per_layer[l][0].numpy(), g["logits"][l]) < 1e-5


def test_fp64_oracle_close_to_fp32_reference():
    g = load_golden("g1500_k2")
    m = oracle_model(g["wseed"], torch.float64)
    _, _, _, per_layer = oracle_forward(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]],
                                        dtype=torch.float64)
    for l in range(3):
        assert rel_err(per_layer[l][0].numpy(), g["logits"][l]) < 1e-3


def test_batched_oracle_equals_per_scene():
    """Eval-mode scenes are independent: a batch of 2 == the two scenes run alone."""
    ga, gb = load_golden("g1500_k2"), load_golden("g2500_k1_5cm")
    m = oracle_model(1)
    cb = gb["coords"].copy()
    cb[:, 0] = 1
    coords = np.concatenate([ga["coords"], cb], 0)
    feats = np.concatenate([ga["feats"], gb["feats"]], 0)
    raw = np.concatenate([ga["raw_coords"], gb["raw_coords"]], 0)
    _, _, _, both = oracle_forward(m, coords, feats, raw, [ga["clicks"], gb["clicks"]], [ga["times"], gb["times"]])
    _, _, _, only_a = oracle_forward(m, ga["coords"], ga["feats"], ga["raw_coords"], [ga["clicks"]], [ga["times"]])
    _, _, _, only_b = oracle_forward(m, gb["coords"], gb["feats"], gb["raw_coords"], [gb["clicks"]], [gb["times"]])
    assert rel_err(both[2][0].numpy(), only_a[2][0].numpy()) < 1e-5
    assert rel_err(both[2][1].numpy(), only_b[2][0].numpy()) < 1e-5
