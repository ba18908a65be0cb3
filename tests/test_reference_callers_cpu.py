"""The reference's OWN training loop (engine.py:26-179, unmodified, imported from /root/reference) driving this package's
model, criterion and optimizer through the alias modules of agile3d_b200.compat: "callers unchanged" demonstrated rather
than asserted.  The C-ABI ops are replaced by their contract emulations (no GPU here); the GPU run of the same loop is
tools/run_reference_engine.py.  Skipped where the reference tree is absent (the GPU box)."""
import importlib
import os
import random
import sys

import numpy as np
import pytest
import torch

import emulate
from helpers import load_golden

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "engine.py")), reason="reference tree not present")


def test_reference_train_one_epoch_runs_unchanged_on_this_package(monkeypatch):
    import agile3d_b200
    from agile3d_b200 import compat
    from agile3d_b200.optim import FlatAdamW
    from agile3d_b200.weights import default_args, synth_state_dict
    emulate.patch_ops(monkeypatch)
    monkeypatch.syspath_prepend(REF)
    saved = {k: sys.modules.get(k) for k in ("MinkowskiEngine", "models", "utils", "utils.seg", "utils.misc", "engine", "evaluation",
                                            "evaluation.evaluator_MO", "matplotlib", "matplotlib.pyplot")}
    try:
        for k in ("utils", "utils.seg", "utils.misc", "engine", "evaluation", "evaluation.evaluator_MO"):
            sys.modules.pop(k, None)
        compat.install()
        monkeypatch.setenv("WANDB_MODE", "disabled")
        engine = importlib.import_module("engine")                      # the reference's file, as is
        assert engine.__file__.startswith(REF)
        g = load_golden("train_g1200_k2")
        args = default_args()
        model = sys.modules["models"].build_model(args)
        model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=g["wseed"]))
        criterion = sys.modules["models"].build_criterion(args)
        optimizer = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
        optimizer.param_groups = [{"lr": 1e-4}]                         # engine.py:161 reads the learning rate for its log
        # one batch in the layout of the reference's collation_fn (datasets/InterMultiObj3DSegDataset.py:126-136)
        coords = torch.from_numpy(g["coords"])
        labels = [torch.from_numpy(g["targets"].astype(np.int64))]
        batch = (coords, torch.from_numpy(g["raw_coords"]), torch.from_numpy(g["feats"]), labels, None, None, [{}], ["scene0"], [2])
        before = model.lin_squeeze_head.kernel.detach().clone()
        random.seed(0)
        np.random.seed(0)
        torch.manual_seed(0)
        monkeypatch.setattr(random, "randint", lambda a, b: 1)         # one pre-sampling click round instead of up to 19
        stats, it = engine.train_one_epoch(model, criterion, [batch], optimizer, torch.device("cpu"), epoch=0, train_total_iter=0,
                                           max_norm=0.1)
        assert it == 1 and np.isfinite(stats["loss"]) and 0.0 <= stats["mIoU"] <= 1.0
        assert float((model.lin_squeeze_head.kernel.detach() - before).abs().max()) > 0      # the optimizer stepped
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
