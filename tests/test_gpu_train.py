"""GPU parity tests (-m gpu) of the training step (e): every training C-ABI entry point
against its fp64 contract emulation (tests/emulate.py — the decoder backward emulations are themselves checked against
torch.autograd in tests/test_host_model_cpu.py), then the whole step (train-mode forward, SetCriterion, backward,
clip + AdamW) against the golden vector of the UNMODIFIED reference in train mode and against the fp64 CPU oracle on a
batch of two scenes.

Tolerances: fp32 kernels vs fp64 emulation 1e-4 .. 1e-3 relative (max|a-b| / max|b|).  The decoder takes DISCRETE
decisions between layers (argmax labels -> attention mask of the next layer, agile3d.py:365-380) and with ~1000 voxels
some voxel always sits within ~1e-4 (relative) of a label boundary, so one rounding difference can flip a voxel and move
losses/gradients of the later layers by ~1e-2 — in the reference as much as here.  The whole-step tests are therefore
tight only on what precedes the first decision (backbone features, BatchNorm statistics, first-layer logits and
losses) and loose (5e-2) on the rest; the gradient arithmetic itself is pinned tightly by the op-level tests above, by
the backbone-only backward test (no discrete decisions: 63 conv/BN layers against the fp64 oracle) and by the CPU
host-logic test against the reference golden (tests/test_host_model_cpu.py).
"""
import json

import numpy as np
import pytest
import torch

import emulate
from helpers import load_golden, oracle_model, oracle_train_step, rel_err
from test_gpu_parity import _decoder_inputs, _random_cloud

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agile3d_b200._lib import lib
    lib()


def t(v):
    return v.to(DEV) if v is not None else None


# ------------------------------------------------------------------------------------------------ BatchNorm
@pytest.mark.parametrize("n,c", [(5000, 32), (777, 96), (33, 256), (120001, 64), (16, 128)])
def test_bn_train_kernels(n, c):
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(n + c)
    z = torch.randn((n, c), generator=g) * 1.7 + 0.4
    res = torch.randn((n, c), generator=g)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    d = lambda v: v.double()
    rm_ref, rv_ref = d(rm).clone(), d(rv).clone()
    mean_r, inv_r = emulate.bn_stats(d(z), 1e-5, 0.02, rm_ref, rv_ref)
    y_r = emulate.bn_apply(d(z), mean_r, inv_r, d(gamma), d(beta), torch.empty(n, c, dtype=torch.float64), d(res), True)
    # z is a channel slice of a wider buffer, y another slice
    zb = torch.zeros((n, c + 32), device=DEV)
    zb[:, 32:] = t(z)
    rm_g, rv_g = t(rm).clone(), t(rv).clone()
    mean, inv = ops.bn_stats(zb[:, 32:], 1e-5, 0.02, rm_g, rv_g)
    assert rel_err(mean.cpu(), mean_r) < 1e-5 and rel_err(inv.cpu(), inv_r) < 1e-4
    assert rel_err(rm_g.cpu(), rm_ref) < 1e-5 and rel_err(rv_g.cpu(), rv_ref) < 1e-4
    yb = torch.full((n, c + 64), -3.0, device=DEV)
    ops.bn_apply(zb[:, 32:], mean, inv, t(gamma), t(beta), yb[:, :c], residual=t(res), relu=True)
    assert rel_err(yb[:, :c].cpu(), y_r) < 1e-4
    assert bool((yb[:, c:] == -3.0).all())
    dy = torch.randn((n, c), generator=g)
    dz_r, g_r = torch.empty(n, c, dtype=torch.float64), torch.empty(n, c, dtype=torch.float64)
    dgam_r, dbet_r = emulate.bn_bwd(d(z), y_r, d(dy), mean_r, inv_r, d(gamma), dz_r, True, g_r)
    dyg = t(dy).clone()
    gout = torch.empty((n, c), device=DEV)
    dgam, dbet = ops.bn_bwd(zb[:, 32:], yb[:, :c], dyg, mean, inv, t(gamma), dyg, relu=True, g_out=gout)   # in place
    scale = max(float(dz_r.abs().max()), 1e-12)
    assert float((dyg.cpu().double() - dz_r).abs().max()) / scale < 2e-3
    assert rel_err(gout.cpu(), g_r) < 1e-6
    assert rel_err(dgam.cpu(), dgam_r) < 1e-3 and rel_err(dbet.cpu(), dbet_r) < 1e-4
    assert rel_err(ops.col_sum(zb[:, 32:]).cpu(), d(z).sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------------ conv gradients
@pytest.mark.parametrize("cin,cout,ks", [(32, 32, 3), (64, 96, 3), (128, 256, 3), (384, 256, 3), (32, 32, 2), (96, 128, 1)])
def test_spconv_bwd_weight_vs_emulation(cin, cout, ks):
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(3000, 28, seed=cin + cout, batch=2))
    n = coords.shape[0]
    g = torch.Generator().manual_seed(cin * 3 + cout)
    x = torch.randn((n, cin), generator=g)
    if ks == 3:
        nbr, n_out = emulate.kernel_map(coords, coords, 0, 3, 1), n
    elif ks == 2:
        coarse, _, _, _ = emulate.downsample(coords, 2)
        nbr, n_out = emulate.kernel_map(coarse, coords, 0, 2, 1), coarse.shape[0]
    else:
        nbr, n_out = None, n
    dout = torch.randn((n_out, cout), generator=g)
    ref = emulate.spconv_bwd_weight(x.double(), nbr, dout.double(), ks ** 3)
    xb = torch.zeros((n, cin + 32), device=DEV)
    xb[:, :cin] = t(x)
    got = ops.spconv_bwd_weight(xb[:, :cin], t(nbr), t(dout), ks ** 3)
    assert rel_err(got.cpu(), ref) < 1e-4
    again = ops.spconv_bwd_weight(xb[:, :cin], t(nbr), t(dout), ks ** 3, dweight=got.clone(), accumulate=True)
    assert rel_err(again.cpu(), 2 * ref) < 1e-4
    assert torch.equal(ops.spconv_bwd_weight(xb[:, :cin], t(nbr), t(dout), ks ** 3), got), "deterministic"


@pytest.mark.parametrize("cin,cout,ks", [(32, 64, 3), (96, 96, 3), (64, 64, 2), (128, 96, 1), (384, 256, 3), (32, 32, 3)])
def test_spconv_bwd_weight_tensor_core_vs_emulation(cin, cout, ks):
    """ag3d_spconv_bwd_weight_tc (tcgen05, both operands as bf16 hi/lo split rows gathered by TMA) against the fp64
    restatement on real kernel maps (3x3x3, stride 2, identity); input rows are a channel slice of a wider buffer."""
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(3000, 28, seed=cin + cout, batch=2))
    n = coords.shape[0]
    g = torch.Generator().manual_seed(cin * 3 + cout)
    x = torch.randn((n, cin), generator=g)
    if ks == 3:
        nbr, n_out = emulate.kernel_map(coords, coords, 0, 3, 1), n
    elif ks == 2:
        coarse, _, _, _ = emulate.downsample(coords, 2)
        nbr, n_out = emulate.kernel_map(coarse, coords, 0, 2, 1), coarse.shape[0]
    else:
        nbr, n_out = None, n
    dout = torch.randn((n_out, cout), generator=g)
    ref = emulate.spconv_bwd_weight(x.double(), nbr, dout.double(), ks ** 3)
    assert ops.wgrad_tc_supported(ks ** 3, cin, cout)
    xb = torch.zeros((n, cin + 32), device=DEV)
    xb[:, :cin] = t(x)
    xs, ds = ops.pack_split_rows(xb[:, :cin]), ops.pack_split_rows(t(dout))
    assert rel_err(ops.unpack_split(xs).cpu(), x) < 1e-5          # split rows carry hi + lo = x up to 2^-17
    got = ops.spconv_bwd_weight_tc(xs, t(nbr), ds, ks ** 3)
    assert rel_err(got.cpu(), ref) < 1e-4
    again = ops.spconv_bwd_weight_tc(xs, t(nbr), ds, ks ** 3, dweight=got.clone(), accumulate=True)
    assert rel_err(again.cpu(), 2 * ref) < 1e-4
    assert torch.equal(ops.spconv_bwd_weight_tc(xs, t(nbr), ds, ks ** 3), got), "deterministic"


@pytest.mark.parametrize("algo", [1, 2], ids=["simt", "tc"])
def test_spconv_bwd_data_is_conv_over_transposed_map(algo):
    """d/dx of sum(conv(x) * dout) from torch.autograd == ag3d_spconv_fwd(dout, transposed map, W^T) for the three map
    kinds the U-Net uses (3x3x3, stride-2, transposed stride-2)."""
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(2500, 26, seed=21, batch=2))
    coarse, _, _, par = emulate.downsample(coords, 2)
    k3 = emulate.kernel_map(coords, coords, 0, 3, 1)
    down = emulate.kernel_map(coarse, coords, 0, 2, 1)
    up = emulate.kernel_map_transposed(coords, par, 1)
    g = torch.Generator().manual_seed(5)
    cases = [("k3", k3, k3, coords.shape[0], coords.shape[0], 27), ("down", down, up, coords.shape[0], coarse.shape[0], 8),
             ("up", up, down, coarse.shape[0], coords.shape[0], 8)]
    for name, nbr, nbr_t, n_in, n_out, K in cases:
        cin, cout = 64, 96
        x = torch.randn((n_in, cin), generator=g, dtype=torch.float64, requires_grad=True)
        w = torch.randn((K, cin, cout), generator=g, dtype=torch.float64) * 0.1
        dout = torch.randn((n_out, cout), generator=g, dtype=torch.float64)
        skip = torch.randn((n_in, cin), generator=g, dtype=torch.float64)
        y = emulate.spconv_fwd(x, nbr, w, torch.empty(n_out, cout, dtype=torch.float64))
        # emulate.spconv_fwd copies into `out`; rebuild the differentiable value
        acc = torch.zeros((n_out, cout), dtype=torch.float64)
        for k in range(K):
            sel = torch.nonzero(nbr[k] >= 0).squeeze(1)
            acc = acc.index_add(0, sel, x[nbr[k][sel].long()] @ w[k])
        (acc * dout).sum().backward()
        wt = (w.flip(0) if K == 27 else w).transpose(1, 2).contiguous().float()
        din = torch.empty((n_in, cin), device=DEV)
        ops.spconv_fwd(t(dout.float()), t(nbr_t), t(wt), din, residual=t(skip.float()), algo=algo,
                       weight_tc=ops.prepare_tc_weight(t(wt)) if algo == 2 else None)
        assert rel_err(din.cpu(), x.grad + skip) < (1e-5 if algo == 1 else 2e-4), name


def test_stem_bwd_weight_vs_emulation():
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(5000, 30, seed=9, batch=2, negative=True))
    g = torch.Generator().manual_seed(1)
    f = torch.rand((coords.shape[0], 3), generator=g)
    dz = torch.randn((coords.shape[0], 32), generator=g)
    ref = emulate.stem_bwd_weight(coords, f.double(), coords, 0, 5, dz.double())
    table, cap, _ = ops.hash_build(t(coords))
    got = ops.stem_bwd_weight(t(coords), t(f), table, cap, 5, t(dz))
    assert rel_err(got.cpu(), ref) < 1e-4
    from agile3d_b200.backbone import CoordinateMaps
    maps = CoordinateMaps(t(coords))          # same neighbours through the bricks of the tensor-stride-4 level
    got_b = ops.stem_bwd_weight(maps.coords[0], t(f), maps.tables[0], maps.caps[0], 5, t(dz), bricks=maps.bricks)
    assert rel_err(got_b.cpu(), ref) < 1e-4


# ------------------------------------------------------------------------------------------------ decoder backward
@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 27, 9), (65, 32, 12)])
def test_c2s_bwd_vs_emulation(nv, nq, n_obj):
    from agile3d_b200 import ops
    H = 8
    g, x, pos, qf, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv + nq)
    label = torch.randint(0, max(n_obj - 1, 1), (nv,), generator=g).to(torch.uint8)    # last object never occurs
    cnt = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    for lab in (None, label):
        lse_g = torch.empty(H * nq, device=DEV)
        ctx_g = ops.c2s_attn_fwd(t(x), t(pos), t(qf), nq, H, t(lab), t(q_obj) if lab is not None else None,
                                 t(cnt) if lab is not None else None, lse=lse_g)
        d = lambda v: v.double()
        lse_r = torch.empty(H * nq, dtype=torch.float64)
        ctx_r = emulate.c2s_attn_fwd(d(x), d(pos), d(qf), nq, H, lab, q_obj, cnt, lse=lse_r)
        assert rel_err(lse_g.cpu(), lse_r) < 1e-4
        dctx = torch.randn((H * nq, 128), generator=g)
        hqp = ops.decoder_bwd_rows(nq, H)
        pad = lambda v, r: torch.cat([v, torch.zeros((r - v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype)])
        ro = torch.full((H * nq,), -1, dtype=torch.int32) if lab is None else \
            torch.where(cnt[q_obj.long()] > 0, q_obj, torch.full_like(q_obj, -1)).repeat(H)
        rowobj = torch.cat([ro, torch.full((hqp - H * nq,), -2, dtype=torch.int32)])
        qp, dp = pad(qf, hqp), pad(dctx, hqp)
        lse_p = torch.cat([lse_r, torch.full((hqp - H * nq,), float("inf"), dtype=torch.float64)])
        dr = pad((d(dctx) * ctx_r).sum(1), hqp)
        dx_r, ds_r = emulate.c2s_attn_bwd(d(x), d(pos), d(qp), d(qp).T.contiguous(), d(dp), d(dp).T.contiguous(), lse_p,
                                          dr, rowobj, hqp, lab)
        dx, ds = ops.c2s_attn_bwd(t(x), t(pos), t(qp), t(qp.T.contiguous()), t(dp), t(dp.T.contiguous()),
                                  t(lse_p.float()), t(dr.float()), t(rowobj), hqp, t(lab))
        assert rel_err(dx.cpu(), dx_r) < 2e-4
        assert rel_err(ds.cpu(), ds_r) < 2e-4
        dq = ops.spconv_bwd_weight(ds, None, t(x), 1)
        ops.spconv_bwd_weight(ds, None, t(pos), 1, dweight=dq, accumulate=True)
        assert rel_err(dq[0].cpu(), ds_r.T @ (d(x) + d(pos))) < 2e-4
        # the same backward with its four GEMMs as 1x1 tcgen05 convolutions (bf16x3) + the point-wise kernel
        dx_t, ds_t = ops.c2s_attn_bwd_tc(t(x), t(pos), t(qp), t(dp), t(lse_p.float()), t(dr.float()), t(rowobj), hqp, t(lab))
        assert rel_err(dx_t.cpu(), dx_r) < 1e-3
        assert rel_err(torch.cat(ds_t, 1).cpu(), ds_r) < 1e-3


@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 25, 9), (129, 32, 12)])
def test_s2c_bwd_vs_emulation(nv, nq, n_obj):
    from agile3d_b200 import ops
    H = 8
    g, x, pos, _, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv * 3 + nq)
    A = torch.randn((H * nq, 128), generator=g) * 0.05
    c = torch.randn(H * nq, generator=g) * 0.1
    U = torch.randn((H * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    dxo, dlg = torch.randn((nv, 128), generator=g), torch.randn((nv, n_obj), generator=g)
    hqp = ops.decoder_bwd_rows(nq, H)
    pad = lambda v, r: torch.cat([v, torch.zeros((r - v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype)])
    Ap, cp, Up, Ep = pad(A, hqp), pad(c, hqp), pad(U, hqp), pad(E, 32)
    d = lambda v: v.double()
    ref = emulate.s2c_mask_bwd(d(x), d(pos), d(Ap), d(Ap).T.contiguous(), d(cp), d(Up), d(Up).T.contiguous(), d(bo), d(lw),
                               d(lb), 1e-5, d(Ep), d(Ep).T.contiguous(), q_obj, nq, H, n_obj, hqp, d(dxo), d(dlg))
    for use_dxo in (True, False):
        got = ops.s2c_mask_bwd(t(x), t(pos), t(Ap), t(Ap.T.contiguous()), t(cp), t(Up), t(Up.T.contiguous()), t(bo), t(lw),
                               t(lb), 1e-5, t(Ep), t(Ep.T.contiguous()), t(q_obj), nq, H, n_obj, hqp,
                               t(dxo) if use_dxo else None, t(dlg))
        if not use_dxo:
            ref = emulate.s2c_mask_bwd(d(x), d(pos), d(Ap), d(Ap).T.contiguous(), d(cp), d(Up), d(Up).T.contiguous(), d(bo),
                                       d(lw), d(lb), 1e-5, d(Ep), d(Ep).T.contiguous(), q_obj, nq, H, n_obj, hqp, None,
                                       d(dlg))
        # the routing of a logit gradient can flip where two queries of one object tie to fp32 noise: compare the rows
        # whose routing agrees (all but a handful)
        same = (got[4].cpu().double() != 0).eq(ref[4] != 0).all(1)
        assert float(same.float().mean()) > 0.995
        for name, a_, b_ in zip(("dx", "a", "ds", "dy", "g"), got[:5], ref[:5]):
            assert rel_err(a_.cpu()[same], b_[same]) < 3e-4, name
        if bool(same.all()):
            assert rel_err(got[5].cpu(), ref[5]) < 1e-3


@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 20, 6), (4100, 25, 9), (129, 32, 12), (3000, 45, 7), (2500, 100, 11),
                                         (6000, 210, 11), (700, 256, 5)])
def test_s2c_bwd_tensor_core_any_query_count(nv, nq, n_obj):
    """ops.s2c_mask_bwd_tc_any (1x1 tcgen05 convolutions + row-wise kernels, heads processed in chunks of <= 256 columns)
    against torch.autograd on the fp64 emulation of the forward, up to 256 click queries per scene."""
    from agile3d_b200 import ops
    H = 8
    g, x, pos, _, _ = _decoder_inputs(nv, min(nq, 40), n_obj, seed=nv * 3 + nq)
    q_obj = torch.randint(0, n_obj, (nq,), generator=g, dtype=torch.int32)
    q_obj[:n_obj] = torch.arange(n_obj, dtype=torch.int32)
    A = torch.randn((H * nq, 128), generator=g) * 0.05
    c = torch.randn(H * nq, generator=g) * 0.1
    U = torch.randn((H * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    dxo, dlg = torch.randn((nv, 128), generator=g), torch.randn((nv, n_obj), generator=g)
    d = lambda v: v.double()
    ref = emulate.s2c_mask_bwd_tc_any(d(x), d(pos), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj, nq, H, n_obj,
                                      d(dxo), d(dlg), None)
    xo = emulate.s2c_mask_fwd(d(x), d(pos), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj, nq, H, n_obj)[0].float()
    xt_dy = lambda xs, dy: sum(ops.spconv_bwd_weight(v, None, dy, 1) for v in xs)
    got = ops.s2c_mask_bwd_tc_any(t(x), t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, H, n_obj,
                                  t(dxo), t(dlg), t(xo), xt_dy)
    # the routing of a logit gradient can flip where two queries of one object tie to rounding: a handful of voxels then
    # contribute to another query's row of dE; everything else is compared at 2e-3 (bf16x3 GEMMs chained six deep)
    for name, a_, b_, tol in zip(("dx", "dA", "dc", "dU", "dbo", "dln_w", "dln_b", "dE"), got, ref,
                                 (2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 2e-2)):
        assert rel_err(a_.cpu(), b_) < tol, (name, rel_err(a_.cpu(), b_))


# ------------------------------------------------------------------------------------------------ loss / optimizer
@pytest.mark.parametrize("n,C", [(5000, 2), (150001, 6), (300, 11), (1, 3)])
def test_loss_kernels_vs_oracle(n, C):
    from agile3d_b200 import ops
    from oracle import criterion_ref as CR
    g = torch.Generator().manual_seed(n + C)
    logits = (torch.randn((n, C), generator=g) * 3).double().requires_grad_()
    target = torch.randint(0, C, (n,), generator=g)
    w = torch.rand(n, generator=g).double() + 0.5
    ld = CR.criterion({"pred_masks": [logits]}, [target], [w])
    coef = torch.tensor([0.7, 1.9], dtype=torch.float64)
    (ld["loss_bce"] * coef[0] + ld["loss_dice"] * coef[1]).backward()
    sums = ops.loss_fwd(t(logits.detach().float()), t(target.int()), t(w.float()))
    assert abs(float(sums[0]) / n - float(ld["loss_bce"])) < 1e-4 * max(1.0, float(ld["loss_bce"]))
    assert abs(float(sums[1]) / n - float(ld["loss_dice"])) < 1e-4
    dl = ops.loss_bwd(t(logits.detach().float()), t(target.int()), t(w.float()), t(coef.float()))
    assert rel_err(dl.cpu(), logits.grad) < 1e-4
    # a label outside [0, C) (the reference asserts the range; -1 is the dataset's ignore id) must fail loudly: NaN
    if n >= 6:
        bad = target.int().clone()
        bad[n // 2] = -1
        bad[n // 3] = C
        assert bool(torch.isnan(ops.loss_fwd(t(logits.detach().float()), t(bad), t(w.float()))).all())
        dlb = ops.loss_bwd(t(logits.detach().float()), t(bad), t(w.float()), t(coef.float()))
        assert bool(torch.isnan(dlb[n // 2]).all()) and bool(torch.isnan(dlb[n // 3]).all()) and not bool(torch.isnan(dlb[0]).any())


def test_click_loss_weights_vs_oracle():
    from agile3d_b200 import ops
    from oracle import criterion_ref as CR
    g = torch.Generator().manual_seed(2)
    xyz = torch.rand((20000, 3), generator=g) * torch.tensor([8.0, 6.0, 3.0])
    rows = [5, 77, 19999, 1234, 0]
    got = ops.click_loss_weights(t(xyz), t(xyz[rows].contiguous()))
    assert rel_err(got.cpu(), CR.click_loss_weights(xyz.double(), rows)) < 1e-5


def test_flat_adamw_and_clip_vs_torch():
    from agile3d_b200.optim import FlatAdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(257, 33, device=DEV)), torch.nn.Parameter(torch.randn(1001, device=DEV))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ours = FlatAdamW(ps, lr=1e-2, weight_decay=1e-2, max_norm=0.1)
    ref = torch.optim.AdamW(qs, lr=1e-2, weight_decay=1e-2)
    for it in range(4):
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
            assert torch.allclose(p, q, atol=2e-6)


# ------------------------------------------------------------------------------------------------ the whole step
def _gpu_train_model(wseed, algo=None):
    import agile3d_b200
    from agile3d_b200.weights import default_args, synth_state_dict
    m = agile3d_b200.build_model(default_args())
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=wseed))
    m = m.to(DEV).train()
    if algo is not None:
        m.backbone.algo = algo
    return m


def _gpu_train_step(m, coords, feats, raw, clicks, times, targets):
    import agile3d_b200
    from agile3d_b200.weights import default_args
    criterion = agile3d_b200.build_criterion(default_args())
    m.zero_grad()
    x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device=DEV)
    rawg = torch.as_tensor(raw).to(DEV)
    h = m.forward_backbone(x, rawg)
    out = m.forward_mask(*h, clicks, times)
    tg = [torch.as_tensor(v).to(DEV) for v in targets]
    weights = agile3d_b200.cal_click_loss_weights(x.C[:, 0], rawg, torch.cat(tg), clicks)
    loss_dict = criterion(out, tg, weights)
    total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
    total.backward()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    return loss_dict, total, grads, out


@pytest.mark.parametrize("algo", [1, 0], ids=["fp32", "tensor-core"])
def test_train_step_vs_reference_golden(algo):
    """Losses and the gradients of all 268 parameters against the UNMODIFIED reference in train mode."""
    g = load_golden("train_g1200_k2")
    names = json.loads(str(g["grad_names"]))
    loss_names = json.loads(str(g["loss_names"]))
    m = _gpu_train_model(g["wseed"], algo)
    loss_dict, total, grads, out = _gpu_train_step(m, g["coords"], g["feats"], g["raw_coords"], g["clicks"], g["times"],
                                                   [g["targets"]])
    tight = 1e-4 if algo == 1 else 1e-3     # before the first discrete decision (levels under 256 rows run in exact fp32)
    loose = 5e-2                            # after it (see the module docstring)
    ref = dict(zip(loss_names, g["loss_values"]))
    for key in ("loss_bce_0", "loss_dice_0"):
        assert abs(float(loss_dict[key].detach()) - ref[key]) < tight * 10, key
    assert rel_err(m.backbone.bn0.bn.running_mean.cpu().numpy(), g["bn0_running_mean"]) < 1e-4
    got = np.array([float(loss_dict[k_].detach()) for k_ in loss_names])
    assert np.abs(got - g["loss_values"]).max() < loose
    assert sorted(grads) == sorted(names)
    gn = np.array([float(grads[n].double().norm()) for n in names])
    assert abs(np.sqrt((gn ** 2).sum()) - float(g["grad_total_norm"])) / float(g["grad_total_norm"]) < 2 * loose
    assert np.abs(gn - g["grad_norms"]).max() / g["grad_norms"].max() < 2 * loose
    assert rel_err(out["pred_masks"][0].detach().cpu().numpy()[::4], g["logits_last"]) < loose


def _two_scenes():
    from agile3d_b200.scenes import make_clicks, make_scene
    scs, clicks, times, targets = [], [], [], []
    for s in (dict(n=1300, seed=40, k=2, cpo=2, bg=1), dict(n=900, seed=41, k=1, cpo=3, bg=0)):
        sc = make_scene(s["n"], 0.02, seed=s["seed"], n_box=5)
        c, tm, lab = make_clicks(sc, s["k"], s["cpo"], s["bg"], seed=s["seed"])
        scs.append(sc); clicks.append(c); times.append(tm); targets.append(np.minimum(lab, len(c) - 1).astype(np.int32))
    coords = np.concatenate([np.concatenate([np.full((sc["coords"].shape[0], 1), b, np.int32), sc["coords"]], 1)
                             for b, sc in enumerate(scs)], 0)
    feats = np.concatenate([sc["feats"] for sc in scs], 0)
    raw = np.concatenate([sc["raw_coords"] for sc in scs], 0)
    return coords, feats, raw, clicks, times, targets


def _grad_errors(grads, rgrads):
    """(relative L2 error of all gradients as one vector, worst per-parameter max error with a 1 % floor, its name)"""
    num = np.sqrt(sum(float((grads[n].double().cpu() - r.double()).norm()) ** 2 for n, r in rgrads.items()))
    den = np.sqrt(sum(float(r.double().norm()) ** 2 for r in rgrads.values()))
    gmax = max(float(v.abs().max()) for v in rgrads.values())
    worst, name = 0.0, ""
    for n, r in rgrads.items():
        e = float((grads[n].double().cpu() - r.double()).abs().max()) / max(float(r.abs().max()), 1e-2 * gmax)
        if e > worst:
            worst, name = e, n
    return num / den, worst, name


@pytest.mark.parametrize("algo", [1, 0], ids=["fp32", "tensor-core"])
def test_backbone_backward_vs_fp64_oracle(algo):
    """No discrete decisions here: train-mode Res16UNet34C + lin_squeeze_head (63 convolutions, 59 batch-statistics
    BatchNorms, residuals, concats) on a batch of two scenes, loss = <features, R> for a fixed random R; the gradients
    of all backbone parameters against torch.autograd on the fp64 CPU oracle."""
    import agile3d_b200
    from oracle import me_ref as ME
    coords, feats, raw, *_ = _two_scenes()
    R = torch.randn((coords.shape[0], 128), generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ref = oracle_model(7, torch.float64).train()
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).double())
    pcd_r, *_ = ref.forward_backbone(x, torch.as_tensor(raw).double())
    (pcd_r.F * R).sum().backward()
    rgrads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    m = _gpu_train_model(7, algo)
    xg = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device=DEV)
    pcd, *_ = m.forward_backbone(xg, torch.as_tensor(raw).to(DEV))
    (pcd.F * R.float().to(DEV)).sum().backward()
    grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    # 63 conv kernels (62 backbone + lin_squeeze_head) + the head bias + 62 BatchNorms x (weight, bias)
    assert sorted(grads) == sorted(rgrads) and len(grads) == 63 + 1 + 2 * 62
    fe = rel_err(pcd.F.detach().cpu().numpy(), pcd_r.F.detach().numpy())
    l2, worst, name = _grad_errors(grads, rgrads)
    # The loss <features, R> with a random R makes every gradient a random-sign sum, so ONE ReLU decision that differs
    # (a pre-activation within rounding of zero) moves all gradients upstream of it by ~5e-3 (L2) / ~1e-1 (worst).
    # Measured with tools/grad_diag2.py on B200: block by block the fp32 path is 2e-7 .. 7e-7 away from the fp64 oracle
    # until block8.0, where a single element (|pre-activation| = 3.5e-6) falls on the other side of the ReLU; the fp32
    # torch oracle shows the same effect on some hosts.  So: tight where no ReLU decision is involved (the head, fed by
    # dL/dpcd and the forward features only), bounded by "a few flips" elsewhere.
    head = {n: g for n, g in grads.items() if n.startswith("lin_squeeze_head.")}
    hl2, hworst, _ = _grad_errors(head, {n: rgrads[n] for n in head})
    head_tol = 1e-4 if algo == 1 else 1e-3
    assert len(head) == 2 and hl2 < head_tol and hworst < 10 * head_tol, (hl2, hworst)
    lim = (1e-4, 3e-2, 3e-1) if algo == 1 else (1e-3, 5e-2, 5e-1)    # features: the north-star 1e-3
    assert fe < lim[0], fe
    assert l2 < lim[1], l2
    assert worst < lim[2], (worst, name)


def test_backbone_backward_block_by_block_fp32():
    """The gradient entering and leaving every BasicBlock against torch.autograd on the fp64 oracle, in backward order.
    Blocks are held to 1e-4 until the first block in which a ReLU decision differs from the oracle's (after that, all
    upstream gradients differ legitimately, see test_backbone_backward_vs_fp64_oracle); at least the first block must
    be reached, and in practice (B200) the first flip sits in the second block processed."""
    import agile3d_b200
    from agile3d_b200.backbone import Res16UNet34C
    from oracle import me_ref as ME
    from oracle.agile3d_ref import RefBasicBlock
    coords, feats, raw, *_ = _two_scenes()
    R = torch.randn((coords.shape[0], 128), generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ref = oracle_model(7, torch.float64).train()
    cap, relus = {}, {}

    def block_hook(name):
        def f(mod, inp, out):
            inp[0].F.retain_grad()
            out.F.retain_grad()
            cap[name] = (inp[0].F, out.F)
        return f

    def relu_hook(name):
        def f(mod, inp, out):
            relus.setdefault(name, []).append(out.F.detach())
        return f

    for n, mod in ref.named_modules():
        if isinstance(mod, RefBasicBlock):
            mod.register_forward_hook(block_hook(n.replace("backbone.", "")))
            mod.relu.register_forward_hook(relu_hook(n.replace("backbone.", "")))
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).double())
    pcd_r, *_ = ref.forward_backbone(x, torch.as_tensor(raw).double())
    (pcd_r.F * R).sum().backward()
    order = []
    for stage in (8, 7, 6, 5, 4, 3, 2, 1):
        nb = len(getattr(ref.backbone, f"block{stage}"))
        order += [f"block{stage}.{b}" for b in reversed(range(nb))]
    rec = []
    orig = Res16UNet34C._block_train_bwd

    def patched(self, blk_rec, dout, W, grads):
        d_in = dout.detach().clone()
        masks = (blk_rec[0]["y"].detach() > 0).cpu(), (blk_rec[1]["y"].detach() > 0).cpu()
        dx = orig(self, blk_rec, dout, W, grads)
        rec.append((d_in.double().cpu(), dx.detach().double().cpu(), masks))
        return dx

    Res16UNet34C._block_train_bwd = patched
    try:
        m = _gpu_train_model(7, 1)
        m.backbone.reorder_rows = False          # this test reads per-block tensors of the backbone: keep the caller's row order
        xg = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device=DEV)
        pcd, *_ = m.forward_backbone(xg, torch.as_tensor(raw).to(DEV))
        (pcd.F * R.float().to(DEV)).sum().backward()
    finally:
        Res16UNet34C._block_train_bwd = orig
    assert len(rec) == len(order) == 23

    def rel(a, b):
        return float((a - b).norm() / max(float(b.norm()), 1e-30))

    tight = 0
    for name, (d_in, dx, masks) in zip(order, rec):
        xin, yout = cap[name]
        flips = sum(int((mg != (mo > 0)).sum()) for mg, mo in zip(masks, relus[name]))
        assert rel(d_in, yout.grad) < 1e-4, (name, "dout", rel(d_in, yout.grad))
        if flips:
            break
        assert rel(dx, xin.grad) < 1e-4, (name, "dx", rel(dx, xin.grad))
        tight += 1
    assert tight >= 1, "a ReLU decision already differs in the last block"


@pytest.mark.parametrize("algo", [1, 0], ids=["fp32", "tensor-core"])
def test_train_step_batch_of_two_vs_fp64_oracle(algo):
    """Whole step on a batch of two scenes (BatchNorm statistics couple them) against the fp64 CPU oracle."""
    coords, feats, raw, clicks, times, targets = _two_scenes()
    ref_m = oracle_model(7, torch.float64)
    rl, rtotal, rgrads, _, rout = oracle_train_step(ref_m, coords, feats, raw, clicks, times, targets, torch.float64)
    m = _gpu_train_model(7, algo)
    loss_dict, total, grads, out = _gpu_train_step(m, coords, feats, raw, clicks, times, targets)
    tight = 1e-4 if algo == 1 else 5e-3
    for b in range(2):                                   # first decoder layer: no mask yet, no discrete decision
        e = rel_err(out["aux_outputs"][0]["pred_masks"][b].detach().cpu().numpy(),
                    rout["aux_outputs"][0]["pred_masks"][b].detach().numpy())
        assert e < tight, (b, e)
    for key in ("loss_bce_0", "loss_dice_0"):
        assert abs(float(loss_dict[key].detach()) - float(rl[key])) < 10 * tight, key
    assert abs(float(total) - float(rtotal)) < 5e-2 * max(1.0, abs(float(rtotal))), (float(total), float(rtotal))
    l2, worst, name = _grad_errors(grads, rgrads)
    assert l2 < 0.15, l2                                  # loose: label flips (module docstring)


def test_training_reduces_the_loss():
    """Five clip + AdamW steps on one small scene: the loss goes down and nothing turns non-finite."""
    import agile3d_b200
    from agile3d_b200.optim import FlatAdamW
    from agile3d_b200.weights import default_args
    g = load_golden("train_g1200_k2")
    m = _gpu_train_model(g["wseed"])
    criterion = agile3d_b200.build_criterion(default_args())
    opt = FlatAdamW(m.parameters(), lr=2e-4, weight_decay=1e-4, max_norm=0.1)
    x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(g["coords"]), features=torch.as_tensor(g["feats"]), device=DEV)
    raw = torch.as_tensor(g["raw_coords"]).to(DEV)
    tg = [torch.as_tensor(g["targets"]).to(DEV)]
    weights = agile3d_b200.cal_click_loss_weights(x.C[:, 0], raw, tg[0], g["clicks"])
    hist = []
    for _ in range(6):
        opt.zero_grad()
        out = m.forward_mask(*m.forward_backbone(x, raw), g["clicks"], g["times"])
        ld = criterion(out, tg, weights)
        total = sum(ld[k] * criterion.weight_dict[k] for k in ld)
        total.backward()
        norm = opt.step()
        assert torch.isfinite(total) and torch.isfinite(norm)
        hist.append(float(total))
    # the loss of this tiny scene is not monotone (label decisions flip between steps, the stem weight gradient sums
    # with shared-memory atomics, so runs differ in the last bits): it must have dropped clearly at some later step
    assert min(hist[2:]) < 0.9 * hist[0], hist
