"""GPU parity tests (-m gpu): every C-ABI entry point against the CPU oracle on the same seeded inputs, the golden
vectors of the unmodified reference end to end, and size-independent properties at BASELINE.json's full size.

Tolerances: integer/index outputs bit-exact; fp32 features and mask logits within 1e-3 relative
(max|a-b| / max|b|, BASELINE.json north_star) — op-level checks use 1e-4 or tighter where the arithmetic is fp32.
"""
import numpy as np
import pytest
import torch

import emulate
from helpers import GOLDEN_CASES, decision_forced_errors, load_golden, oracle_forward, oracle_model, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _cuda_ok():
    return torch.cuda.is_available()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not _cuda_ok():
        pytest.skip("no CUDA device")
    from agile3d_b200._lib import lib
    lib()     # raises if the library is not built: GPU tests must never pass on a fallback


def _random_cloud(n, extent, seed, batch=1, negative=False):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(batch):
        c = rng.integers(-extent if negative else 0, extent, size=(n, 3))
        # surface-like: squash one axis
        c[:, 2] = c[:, 2] // 8
        c = np.unique(c, axis=0)
        rng.shuffle(c)
        rows.append(np.concatenate([np.full((c.shape[0], 1), b), c], 1))
    return np.concatenate(rows, 0).astype(np.int32)


# ------------------------------------------------------------------------------------------------ coordinates
@pytest.mark.parametrize("n,extent,batch,neg", [(500, 20, 1, False), (20000, 64, 2, True), (3000, 40, 3, False)])
def test_maps_bit_exact(n, extent, batch, neg):
    from agile3d_b200.backbone import CoordinateMaps
    coords = _random_cloud(n, extent, seed=n, batch=batch, negative=neg)
    maps = CoordinateMaps(torch.from_numpy(coords).to(DEV), count_pairs=True)
    # oracle
    cur = torch.from_numpy(coords)
    levels, parents = [cur], []
    for lvl in range(4):
        c, _, _, par = emulate.downsample(levels[-1], 2 << lvl)
        levels.append(c)
        parents.append(par)
    for lvl in range(5):
        assert torch.equal(maps.coords[lvl].cpu(), levels[lvl]), f"coarse coords level {lvl}"
        ref = emulate.kernel_map(levels[lvl], levels[lvl], 0, 3, 1 << lvl)
        assert torch.equal(maps.k3[lvl].cpu(), ref), f"3x3x3 map level {lvl}"
        assert maps.pair_counts[("k3", lvl)] == int((ref >= 0).sum())
    for lvl in range(4):
        assert torch.equal(maps.parents[lvl].cpu(), parents[lvl]), f"parents level {lvl}"
        ref = emulate.kernel_map(levels[lvl + 1], levels[lvl], 0, 2, 1 << lvl)
        assert torch.equal(maps.down[lvl].cpu(), ref), f"stride-2 map level {lvl}"
        ref_t = emulate.kernel_map_transposed(levels[lvl], parents[lvl], 1 << lvl)
        assert torch.equal(maps.up[lvl].cpu(), ref_t), f"transposed map level {lvl}"


@pytest.mark.parametrize("n,extent,batch", [(20000, 64, 2), (3000, 40, 3)])
def test_internal_row_order_is_a_relabelling(n, extent, batch):
    """CoordinateMaps(reorder=True): the permutation equals the stable (scene, neighbour pattern) sort of the contract
    emulation, and every neighbour table is the canonical one rewritten into the new numbering - which is how kernel-map
    parity is stated once rows are re-ordered internally (compare after un-permuting)."""
    from agile3d_b200.backbone import CoordinateMaps
    coords = torch.from_numpy(_random_cloud(n, extent, seed=n + 1, batch=batch)).to(DEV)
    ref = CoordinateMaps(coords)
    CoordinateMaps.REORDER_LEVELS = (0, 1, 2)          # exercise the coarse levels too (the model re-orders level 0 only)
    try:
        got = CoordinateMaps(coords, reorder=True)
    finally:
        CoordinateMaps.REORDER_LEVELS = (0,)
    for l in range(5):
        if got.perm[l] is None:
            assert l not in (0, 1, 2) or ref.sizes[l] < 256
            assert torch.equal(got.coords[l], ref.coords[l])
            continue
        perm_ref, inv_ref = emulate.row_order(ref.k3[l].cpu(), ref.coords[l].cpu())
        assert torch.equal(got.perm[l].cpu(), perm_ref) and torch.equal(got.inv[l].cpu(), inv_ref)
        assert torch.equal(got.coords[l], ref.coords[l][got.perm[l].long()])
        b = got.coords[l][:, 0]
        assert bool((b[1:] >= b[:-1]).all())                                   # scenes stay contiguous and ordered
    ident = lambda m: torch.arange(m, dtype=torch.int32, device=DEV)
    P = [got.perm[l] if got.perm[l] is not None else ident(ref.sizes[l]) for l in range(5)]
    I = [got.inv[l] if got.inv[l] is not None else ident(ref.sizes[l]) for l in range(5)]
    relabel = lambda nbr, p_out, i_in: torch.where(nbr[:, p_out.long()] >= 0, i_in[nbr[:, p_out.long()].clamp(min=0).long()],
                                                   nbr[:, p_out.long()])
    for l in range(5):
        assert torch.equal(got.k3[l], relabel(ref.k3[l], P[l], I[l])), f"k3 level {l}"
    for l in range(4):
        assert torch.equal(got.down[l], relabel(ref.down[l], P[l + 1], I[l])), f"down {l}"
        assert torch.equal(got.up[l], relabel(ref.up[l], P[l], I[l + 1])), f"up {l}"


def test_row_order_changes_nothing_for_the_caller():
    """same scene with and without the internal row order: features, encodings and logits in the caller's order agree to
    fp32 reduction-order noise (the order only changes which rows share a tile)"""
    g = load_golden("g3000_k3")
    m = _gpu_model(g["wseed"])
    h1, l1 = _run_gpu(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    m.backbone.reorder_rows = False
    h0, l0 = _run_gpu(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    assert h1[0].perm is not None and h0[0].perm is None
    assert rel_err(h1[0].F.cpu().numpy(), h0[0].F.cpu().numpy()) < 2e-5
    assert torch.equal(h1[3][4][0][0], h0[3][4][0][0])
    for l in range(3):
        assert rel_err(l1[l][0].cpu().numpy(), l0[l][0].cpu().numpy()) < 1e-4


def test_kernel_map_5x5x5_and_dilation():
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(4000, 30, seed=11))
    table, cap, _ = ops.hash_build(coords.to(DEV))
    got = ops.kernel_map(coords.to(DEV), table, cap, 5, 1)
    assert torch.equal(got.cpu(), emulate.kernel_map(coords, coords, 0, 5, 1))
    got = ops.kernel_map(coords.to(DEV), table, cap, 3, 1, dilation=2)
    assert torch.equal(got.cpu(), emulate.kernel_map(coords, coords, 0, 3, 1, 2))


def test_duplicate_and_out_of_range_coordinates_rejected():
    from agile3d_b200.backbone import CoordinateMaps
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3], [0, 4, 4, 4]], dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        CoordinateMaps(c)
    c = torch.tensor([[0, 1, 2, 3], [0, 40000, 2, 3]], dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        CoordinateMaps(c)


def test_single_voxel_scene():
    from agile3d_b200.backbone import CoordinateMaps
    m = CoordinateMaps(torch.tensor([[0, 5, 5, 5]], dtype=torch.int32, device=DEV))
    assert m.sizes == [1, 1, 1, 1, 1]
    assert m.k3[0].cpu()[:, 0].tolist() == [-1] * 13 + [0] + [-1] * 13


# ------------------------------------------------------------------------------------------------ sparse conv
@pytest.mark.parametrize("cin,cout,ks", [(32, 32, 3), (64, 96, 3), (96, 128, 3), (128, 256, 3), (384, 256, 3),
                                         (32, 32, 2), (256, 128, 2)])
def test_spconv_simt_vs_oracle(cin, cout, ks):
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(3000, 28, seed=cin + cout, batch=2))
    n = coords.shape[0]
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn((n, cin), generator=g)
    w = torch.randn((ks ** 3, cin, cout), generator=g) / np.sqrt(cin * 8)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    if ks == 3:
        nbr = emulate.kernel_map(coords, coords, 0, 3, 1)
        out_n = n
    else:
        coarse, _, _, _ = emulate.downsample(coords, 2)
        nbr = emulate.kernel_map(coarse, coords, 0, 2, 1)
        out_n = coarse.shape[0]
    res = torch.randn((out_n, cout), generator=g)
    ref = emulate.spconv_fwd(x, nbr, w, torch.empty(out_n, cout), scale, shift, res, relu=True)
    # input is a channel slice of a wider buffer, output a slice of another: exercises the leading dimensions
    xin = torch.zeros((n, cin + 32), device=DEV)
    xin[:, 32:] = x.to(DEV)
    outb = torch.full((out_n, cout + 64), -7.0, device=DEV)
    ops.spconv_fwd(xin[:, 32:], nbr.to(DEV), w.to(DEV), outb[:, 64:], scale.to(DEV), shift.to(DEV), res.to(DEV),
                   relu=True, algo=ops.ALGO_SIMT)
    assert rel_err(outb[:, 64:].cpu().numpy(), ref.numpy()) < 1e-5
    assert bool((outb[:, :64] == -7.0).all()), "wrote outside the channel slice"


def test_spconv_1x1_and_transposed():
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(2500, 26, seed=5))
    coarse, _, _, par = emulate.downsample(coords, 2)
    g = torch.Generator().manual_seed(3)
    xc = torch.randn((coarse.shape[0], 64), generator=g)
    w = torch.randn((8, 64, 96), generator=g) * 0.1
    nbr_t = emulate.kernel_map_transposed(coords, par, 1)
    ref = emulate.spconv_fwd(xc, nbr_t, w, torch.empty(coords.shape[0], 96))
    got = torch.empty((coords.shape[0], 96), device=DEV)
    ops.spconv_fwd(xc.to(DEV), nbr_t.to(DEV), w.to(DEV), got, algo=ops.ALGO_SIMT)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-5
    w1 = torch.randn((96, 128), generator=g) * 0.1
    b1 = torch.randn(128, generator=g)
    ref1 = ref @ w1 + b1
    got1 = torch.empty((coords.shape[0], 128), device=DEV)
    ops.spconv_fwd(got, None, w1.to(DEV), got1, None, b1.to(DEV), algo=ops.ALGO_SIMT)
    assert rel_err(got1.cpu().numpy(), ref1.numpy()) < 1e-5


def test_stem_conv_vs_oracle():
    from agile3d_b200 import ops
    coords = torch.from_numpy(_random_cloud(5000, 30, seed=9, batch=2, negative=True))
    g = torch.Generator().manual_seed(1)
    f = torch.rand((coords.shape[0], 3), generator=g)
    w = torch.randn((125, 3, 32), generator=g) * 0.1
    sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    ref = emulate.stem_conv_fwd(coords, f, coords, 0, 5, w, torch.empty(coords.shape[0], 32), sc, sh, relu=True)
    table, cap, _ = ops.hash_build(coords.to(DEV))
    buf = torch.zeros((coords.shape[0], 128), device=DEV)
    ops.stem_conv_fwd(coords.to(DEV), f.to(DEV), table, cap, 5, w.to(DEV), buf[:, 96:], sc.to(DEV), sh.to(DEV))
    assert rel_err(buf[:, 96:].cpu().numpy(), ref.numpy()) < 1e-5
    assert float(buf[:, :96].abs().max()) == 0.0


@pytest.mark.parametrize("ksize", [3, 5])
def test_stem_conv_bricks_identical_to_hash_probes(ksize):
    """the brick lookup (8 probes of the tensor-stride-4 table + the bricks' row lists) finds exactly the neighbours the
    ksize^3 probes of the full-resolution table find: bit-identical stem output, brick table bit-exact vs the emulation"""
    from agile3d_b200 import ops
    from agile3d_b200.backbone import CoordinateMaps
    coords = torch.from_numpy(_random_cloud(6000, 40, seed=21, batch=3, negative=True))
    maps = CoordinateMaps(coords.to(DEV))
    t2, cap2, rows = maps.bricks
    lv, par = [coords], []
    for lvl in range(2):
        c, _, _, p = emulate.downsample(lv[-1], 2 << lvl)
        lv.append(c)
        par.append(p)
    assert torch.equal(rows.cpu(), emulate.brick_rows(coords, par[0], par[1], lv[2].shape[0]))
    g = torch.Generator().manual_seed(2)
    f = torch.rand((coords.shape[0], 3), generator=g).to(DEV)
    w = (torch.randn((ksize ** 3, 3, 32), generator=g) * 0.1).to(DEV)
    a = torch.empty((coords.shape[0], 32), device=DEV)
    b = torch.empty_like(a)
    ops.stem_conv_fwd(maps.coords[0], f, maps.tables[0], maps.caps[0], ksize, w, a, relu=False)
    ops.stem_conv_fwd(maps.coords[0], f, maps.tables[0], maps.caps[0], ksize, w, b, relu=False, bricks=maps.bricks)
    assert torch.equal(a, b)
    ref = emulate.stem_conv_fwd(coords, f.cpu(), coords, 0, ksize, w.cpu(), torch.empty(coords.shape[0], 32), relu=False)
    assert rel_err(b.cpu().numpy(), ref.numpy()) < 1e-5
    # rows in the internal (re-ordered) order still read the features in the caller's order
    maps._reorder()
    if maps.perm[0] is not None:
        c = torch.empty_like(a)
        ops.stem_conv_fwd(maps.coords[0], f, maps.tables[0], maps.caps[0], ksize, w, c, relu=False, bricks=maps.bricks)
        assert torch.equal(c, a[maps.perm[0].long()])


def test_gather_rows_and_split_posenc():
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(8)
    idx = torch.randperm(5003, generator=g).to(torch.int32)
    for cols in (3, 4, 7, 128):                    # 12-byte rows (xyz), 16-byte rows (coordinates), odd, feature rows
        src = torch.randn((5003, cols), generator=g)
        assert torch.equal(ops.gather_rows(src.to(DEV), idx.to(DEV)).cpu(), src[idx.long()])
    ints = torch.randint(-50, 50, (5003, 4), generator=g, dtype=torch.int32)
    assert torch.equal(ops.gather_rows(ints.to(DEV), idx.to(DEV)).cpu(), ints[idx.long()])
    # the split-row output of the positional encoding holds the fp32 output to 2^-17
    xyz = torch.rand((4001, 3), generator=g) * torch.tensor([8.0, 6.0, 3.0])
    B = torch.randn((3, 64), generator=g)
    pos, rng, pos_s = ops.fourier_posenc(xyz.to(DEV), [0, 1500, 4001], B.to(DEV), want_split=True)
    pos0, rng0 = ops.fourier_posenc(xyz.to(DEV), [0, 1500, 4001], B.to(DEV))
    assert torch.equal(pos, pos0) and torch.equal(rng, rng0)
    assert float((ops.unpack_split(pos_s) - pos).abs().max()) < 2e-5
    assert torch.equal(pos_s.view(torch.int32), ops.pack_split_rows(pos).view(torch.int32))      # same bits


# ------------------------------------------------------------------------------------------------ pos-enc
def test_fourier_posenc_vs_oracle():
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(4)
    xyz = torch.rand((7000, 3), generator=g) * torch.tensor([8.0, 6.0, 3.0])
    B = torch.randn((3, 64), generator=g)
    offs = [0, 2500, 7000]
    ref, ref_rng = emulate.fourier_posenc(xyz, offs, B)
    got, rng = ops.fourier_posenc(xyz.to(DEV), offs, B.to(DEV))
    assert torch.equal(rng.cpu(), ref_rng)
    assert float((got.cpu() - ref).abs().max()) < 2e-5


# ------------------------------------------------------------------------------------------------ decoder kernels
def _decoder_inputs(nv, nq, n_obj, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((nv, 128), generator=g)
    pos = torch.randn((nv, 128), generator=g) * 0.7
    qf = torch.randn((8 * nq, 128), generator=g) * 0.08
    n_fg = nq - 10
    q_obj = torch.tensor(sorted((i % (n_obj - 1)) + 1 for i in range(n_fg)) + [0] * 10, dtype=torch.int32) \
        if n_obj > 1 else torch.zeros(nq, dtype=torch.int32)
    return g, x, pos, qf, q_obj


@pytest.mark.parametrize("algo", [1, 2], ids=["simt", "tc"])
@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 27, 9), (3000, 45, 11),
                                         (64, 16, 3), (65, 12, 2)])
def test_c2s_vs_oracle(nv, nq, n_obj, algo):
    from agile3d_b200 import ops
    g, x, pos, qf, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv + nq)
    ref = emulate.c2s_attn_fwd(x.double(), pos.double(), qf.double(), nq, 8)
    got = ops.c2s_attn_fwd(x.to(DEV), pos.to(DEV), qf.to(DEV), nq, 8, algo=algo)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-4
    # masked: the last object's label never occurs -> its rows must be un-masked (agile3d.py:369,375)
    label = torch.randint(0, max(n_obj - 1, 1), (nv,), generator=g).to(torch.uint8)
    cnt = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    ref = emulate.c2s_attn_fwd(x.double(), pos.double(), qf.double(), nq, 8, label, q_obj, cnt)
    got = ops.c2s_attn_fwd(x.to(DEV), pos.to(DEV), qf.to(DEV), nq, 8, label.to(DEV), q_obj.to(DEV), cnt.to(DEV),
                           algo=algo)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-4


@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 27, 9), (3000, 45, 11),
                                         (64, 16, 3), (65, 12, 2), (150001, 20, 6), (1, 11, 2), (63, 20, 4)])
def test_c2s_split_rows_vs_oracle(nv, nq, n_obj):
    """the TMA-fed kernel on split rows: compared with the fp64 oracle evaluated on the values the split rows hold
    (hi + lo, 2^-17 relative from x) and with the fp32-row kernel"""
    from agile3d_b200 import ops
    g, x, pos, qf, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv + nq)
    xs, ps = ops.pack_split_rows(x.to(DEV)), ops.pack_split_rows(pos.to(DEV))
    xv, pv = ops.unpack_split(xs).cpu().double(), ops.unpack_split(ps).cpu().double()
    ref = emulate.c2s_attn_fwd(xv, pv, qf.double(), nq, 8)
    got = ops.c2s_attn_fwd(xs, ps, qf.to(DEV), nq, 8, split=True)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-4
    old = ops.c2s_attn_fwd(x.to(DEV), pos.to(DEV), qf.to(DEV), nq, 8, algo=2)
    assert rel_err(got.cpu().numpy(), old.cpu().numpy()) < 1e-4
    label = torch.randint(0, max(n_obj - 1, 1), (nv,), generator=g).to(torch.uint8)
    cnt = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    ref = emulate.c2s_attn_fwd(xv, pv, qf.double(), nq, 8, label, q_obj, cnt)
    lse = torch.empty(8 * nq, device=DEV)
    got = ops.c2s_attn_fwd(xs, ps, qf.to(DEV), nq, 8, label.to(DEV), q_obj.to(DEV), cnt.to(DEV), split=True, lse=lse)
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-4
    assert bool(torch.isfinite(lse[: 8 * nq]).any())


@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 24, 9), (777, 24, 12),
                                         (128, 16, 4), (129, 20, 5), (150001, 20, 6), (1, 11, 2), (127, 20, 4)])
def test_s2c_mask_split_rows_vs_oracle(nv, nq, n_obj):
    """the TMA-fed kernel on split rows vs the fp64 oracle evaluated on the values the split rows hold"""
    from agile3d_b200 import ops
    g, x, pos, _, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv * 3 + nq)
    A = torch.randn((8 * nq, 128), generator=g) * 0.05
    c = torch.randn(8 * nq, generator=g) * 0.1
    U = torch.randn((8 * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    d = lambda t: t.double()
    t = lambda v: v.to(DEV)
    xs, ps = ops.pack_split_rows(t(x)), ops.pack_split_rows(t(pos))
    xv, pv = ops.unpack_split(xs).cpu(), ops.unpack_split(ps).cpu()
    ry, rl, rlab, rcnt = emulate.s2c_mask_fwd(d(xv), d(pv), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj,
                                              nq, 8, n_obj)
    y, lg, lab, cnt = ops.s2c_mask_fwd(xs, ps, t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8, n_obj,
                                       split=True)
    assert rel_err(ops.unpack_split(y).cpu().numpy(), ry.numpy()) < 1e-4
    assert rel_err(lg.cpu().numpy(), rl.numpy()) < 1e-4
    top2 = torch.topk(rl, min(2, n_obj), dim=1)[0]
    safe = (top2[:, 0] - top2[:, -1]) > 1e-3 if n_obj > 1 else torch.ones(nv, dtype=torch.bool)
    assert torch.equal(lab.cpu()[safe], rlab[safe])
    assert int(cnt.sum()) == nv and torch.equal(cnt.cpu(), torch.bincount(lab.cpu().long(), minlength=n_obj).int())
    # in place (x_out aliases x) and without the feature write: same logits
    xd = xs.clone()
    y2, lg2, _, _ = ops.s2c_mask_fwd(xd, ps, t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8, n_obj,
                                     x_out=xd, split=True)
    assert torch.equal(y2.view(torch.int32), y.view(torch.int32)) and torch.equal(lg2, lg)
    y3, lg3, _, _ = ops.s2c_mask_fwd(xs, ps, t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8, n_obj,
                                     split=True, write_x=False)
    assert y3 is None and torch.equal(lg3, lg)


@pytest.mark.parametrize("algo", [1, 2], ids=["simt", "tc"])
@pytest.mark.parametrize("nv,nq,n_obj", [(1000, 11, 2), (5003, 15, 3), (20000, 20, 6), (4100, 25, 9), (777, 32, 12),
                                         (128, 16, 4), (129, 17, 5)])
def test_s2c_mask_vs_oracle(nv, nq, n_obj, algo):
    from agile3d_b200 import ops
    g, x, pos, _, q_obj = _decoder_inputs(nv, nq, n_obj, seed=nv * 3 + nq)
    A = torch.randn((8 * nq, 128), generator=g) * 0.05
    c = torch.randn(8 * nq, generator=g) * 0.1
    U = torch.randn((8 * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    d = lambda t: t.double()
    ry, rl, rlab, rcnt = emulate.s2c_mask_fwd(d(x), d(pos), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj,
                                              nq, 8, n_obj)
    t = lambda v: v.to(DEV)
    y, lg, lab, cnt = ops.s2c_mask_fwd(t(x), t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq,
                                       8, n_obj, algo=algo)
    assert rel_err(y.cpu().numpy(), ry.numpy()) < 1e-4
    assert rel_err(lg.cpu().numpy(), rl.numpy()) < 1e-4
    # labels: exact wherever the top-2 margin of the fp64 logits exceeds the fp32 noise
    top2 = torch.topk(rl, min(2, n_obj), dim=1)[0]
    safe = (top2[:, 0] - top2[:, -1]) > 1e-3 if n_obj > 1 else torch.ones(nv, dtype=torch.bool)
    assert torch.equal(lab.cpu()[safe], rlab[safe])
    assert int(cnt.sum()) == nv and torch.equal(cnt.cpu(), torch.bincount(lab.cpu().long(), minlength=n_obj).int())
    # in-place variant (x_out aliases x) gives the same answer
    xd = t(x).clone()
    y2, lg2, _, _ = ops.s2c_mask_fwd(xd, t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8,
                                     n_obj, x_out=xd, algo=algo)
    assert torch.equal(y2, y) and torch.equal(lg2, lg)


@pytest.mark.parametrize("nv,nq,n_obj", [(3000, 33, 4), (5003, 45, 7), (4100, 57, 5), (9000, 100, 11), (20000, 210, 11),
                                         (777, 256, 12), (129, 64, 3), (2000, 48, 32)])
def test_s2c_mask_many_queries_vs_oracle(nv, nq, n_obj):
    """More than 32 click queries per scene (the tail of eval_multi_obj.py:116-167): query groups of 16 with two-pass
    softmax statistics (csrc/decoder_mq.cu).  One object (the last) has no query: its logit column is -inf."""
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(nv * 5 + nq)
    x = torch.randn((nv, 128), generator=g)
    pos = torch.randn((nv, 128), generator=g) * 0.7
    q_obj = torch.randint(0, n_obj - 1, (nq,), generator=g, dtype=torch.int32)      # NOT sorted: the kernel sorts
    q_obj[:n_obj - 1] = torch.arange(n_obj - 1, dtype=torch.int32)                  # every other object has a query
    A = torch.randn((8 * nq, 128), generator=g) * 0.05
    c = torch.randn(8 * nq, generator=g) * 0.1
    U = torch.randn((8 * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    d = lambda t: t.double()
    ry, rl, rlab, rcnt = emulate.s2c_mask_fwd(d(x), d(pos), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj,
                                              nq, 8, n_obj)
    t = lambda v: v.to(DEV)
    y, lg, lab, cnt = ops.s2c_mask_fwd(t(x), t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq,
                                       8, n_obj)
    assert rel_err(y.cpu().numpy(), ry.numpy()) < 1e-4
    finite = torch.isfinite(rl)
    assert torch.equal(torch.isfinite(lg.cpu()), finite)
    assert rel_err(lg.cpu()[finite].numpy(), rl[finite].numpy()) < 1e-4
    top2 = torch.topk(rl, 2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(lab.cpu()[safe], rlab[safe])
    assert int(cnt.sum()) == nv and torch.equal(cnt.cpu(), torch.bincount(lab.cpu().long(), minlength=n_obj).int())
    xd = t(x).clone()                                                               # in place
    y2, lg2, _, _ = ops.s2c_mask_fwd(xd, t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8,
                                     n_obj, x_out=xd)
    assert torch.equal(y2, y) and torch.equal(lg2, lg)


@pytest.mark.parametrize("B,nq", [(1, 12), (3, 20), (2, 45), (1, 210), (1, 256), (8, 16)])
def test_query_kernels_vs_emulation(B, nq):
    """K11 (csrc/query_ops.cu): query assembly, c2s fold, c2s tail + c2c projections, c2c attention + FFN + s2c folds +
    mask embeddings against the fp64 contract emulation, with the weight blob of a real model."""
    from agile3d_b200 import ops
    m = _gpu_model(9)
    blob = m._layer_blob(1)
    assert blob.numel() == emulate.query_blob_floats()
    g = torch.Generator().manual_seed(B * 1000 + nq)
    rn = lambda *sh: torch.randn(sh, generator=g)
    # ---- query assembly
    nv = 500
    feats, xyz = rn(nv, 128), torch.rand((nv, 3), generator=g) * 5
    rng = torch.tensor([[0.0, 0.0, 0.0, 5.0, 5.0, 5.0]]).repeat(B, 1) + torch.rand((B, 6), generator=g) * 0.1
    src = torch.randint(0, nv, (B * nq,), generator=g, dtype=torch.int32)
    src[::3] = -(torch.randint(0, 10, (len(src[::3]),), generator=g, dtype=torch.int32) + 1)
    tix = torch.randint(0, 200, (B * nq,), generator=g, dtype=torch.int32)
    sor = torch.arange(B, dtype=torch.int32).repeat_interleave(nq)
    gB, ttab, bgf, bgp = m.pos_enc.gauss_B.cpu(), m.time_encode.cpu(), m.bg_query_feat.weight.detach().cpu(), m.bg_query_pos.weight.detach().cpu()
    t = lambda v: v.to(DEV)
    q_ref, qp_ref = emulate.query_init(feats.double(), xyz.double(), rng.double(), src, tix, sor, gB.double(), ttab.double(),
                                       bgf.double(), bgp.double())
    q, qp = ops.query_init(t(feats), t(xyz), t(rng), t(src), t(tix), t(sor), t(gB), t(ttab), t(bgf), t(bgp))
    assert torch.equal(q.cpu(), q_ref.float())
    assert float((qp.cpu() - qp_ref).abs().max()) < 3e-5
    # ---- the three per-layer kernels
    Q, P = rn(B, nq, 128), rn(B, nq, 128) * 0.7
    ctx = rn(B, 8 * nq, 128)
    bd = blob.detach().cpu().double()
    ref_fold = emulate.query_fold_c2s(Q.double(), P.double(), bd, B, nq)
    got_fold = ops.query_fold_c2s(t(Q), t(P), blob, B, nq)
    assert rel_err(got_fold.cpu().numpy(), ref_fold.numpy()) < 1e-5
    ref_a = emulate.query_update_a(ctx.double(), Q.double(), P.double(), bd, B, nq)
    got_a = ops.query_update_a(t(ctx), t(Q), t(P), blob, B, nq)
    for name, a_, b_ in zip(("q1", "qh", "kh", "vh"), got_a, ref_a):
        assert rel_err(a_.cpu().numpy(), b_.numpy()) < 1e-5, name
    ref_b = emulate.query_update_b(*ref_a, P.double(), bd, B, nq)
    got_b = ops.query_update_b(*[v.float().to(DEV) for v in ref_a], t(P), blob, B, nq)
    for name, a_, b_ in zip(("q3", "A", "c", "U", "E"), got_b, ref_b):
        assert rel_err(a_.cpu().numpy(), b_.numpy()) < 2e-5, name


def test_fused_query_path_equals_torch_glue():
    """eval forward_mask with the fused click-query kernels against the same model running the torch glue."""
    g = load_golden("g3000_k3")
    m = _gpu_model(g["wseed"])
    h, layers = _run_gpu(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    m.fused_queries = False
    out = m.forward_mask(*h, [g["clicks"]], [g["times"]])
    ref = [a["pred_masks"] for a in out["aux_outputs"]] + [out["pred_masks"]]
    for l in range(3):
        assert rel_err(layers[l][0].cpu().numpy(), ref[l][0].cpu().numpy()) < 2e-5


# ------------------------------------------------------------------------------------------------ end to end
def _gpu_model(wseed, algo=None):
    import agile3d_b200
    from agile3d_b200.weights import default_args, synth_state_dict
    m = agile3d_b200.build_model(default_args()).eval()
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=wseed))
    m = m.to(DEV)
    if algo is not None:
        m.backbone.algo = algo
    return m


def _run_gpu(m, coords, feats, raw, clicks, times):
    import agile3d_b200
    x = agile3d_b200.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats), device=DEV)
    h = m.forward_backbone(x, torch.as_tensor(raw).to(DEV))
    out = m.forward_mask(*h, clicks, times)
    layers = [a["pred_masks"] for a in out["aux_outputs"]] + [out["pred_masks"]]
    return h, layers


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_end_to_end_vs_reference_golden(name):
    """The CUDA path against the outputs of the UNMODIFIED reference model files (tests/golden)."""
    g = load_golden(name)
    m = _gpu_model(g["wseed"])
    h, layers = _run_gpu(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    assert [a.shape[0] for a in h[1]] == g["level_sizes"].tolist()
    assert rel_err(h[0].F.cpu().numpy()[::4], g["pcd_features"]) < 1e-3
    assert float((h[3][4][0][0].cpu()[::16] - torch.from_numpy(g["pos_enc"])).abs().max()) < 2e-5
    for l in range(3):
        e = rel_err(layers[l][0].cpu().numpy(), g["logits"][l])
        assert e < 1e-3, f"layer {l}: rel err {e}"


def test_end_to_end_vs_fp64_oracle_batched():
    """Batch of two scenes against the fp64 oracle (the truth for the 1e-3 criterion)."""
    from agile3d_b200.scenes import make_clicks, make_scene
    sa = make_scene(6000, 0.02, seed=21, n_box=8)
    sb = make_scene(9000, 0.05, seed=22, n_box=10)
    ca, ta, _ = make_clicks(sa, 3, 2, 1, seed=1)
    cb, tb, _ = make_clicks(sb, 5, 2, 0, seed=2)
    coords = np.concatenate([np.concatenate([np.zeros((sa["coords"].shape[0], 1), np.int32), sa["coords"]], 1),
                             np.concatenate([np.ones((sb["coords"].shape[0], 1), np.int32), sb["coords"]], 1)], 0)
    feats = np.concatenate([sa["feats"], sb["feats"]], 0)
    raw = np.concatenate([sa["raw_coords"], sb["raw_coords"]], 0)
    ref_m = oracle_model(7, torch.float64)
    _, _, _, ref_layers = oracle_forward(ref_m, coords, feats, raw, [ca, cb], [ta, tb], dtype=torch.float64)
    m = _gpu_model(7)
    _, layers = _run_gpu(m, coords, feats, raw, [ca, cb], [ta, tb])
    for l in range(3):
        for b in range(2):
            e = rel_err(layers[l][b].cpu().numpy(), ref_layers[l][b].numpy())
            assert e < 1e-3, f"layer {l} scene {b}: rel err {e}"


def test_many_click_queries_vs_fp64_oracle():
    """The tail of the iterative-click protocol (eval_multi_obj.py:116-167): more than 32 queries per scene.  c2s runs
    its query groups, s2c its many-query kernel (csrc/decoder_mq.cu); same 1e-3 criterion."""
    from agile3d_b200.scenes import make_clicks, make_scene
    sc = make_scene(5000, 0.02, seed=31, n_box=8)
    clicks, times, _ = make_clicks(sc, 4, 11, 3, seed=5)                     # 44 fg + 3 bg clicks + 10 learned = 57 queries
    coords = np.concatenate([np.zeros((sc["coords"].shape[0], 1), np.int32), sc["coords"]], 1)
    ref_m = oracle_model(7, torch.float64)
    _, _, _, ref_layers = oracle_forward(ref_m, coords, sc["feats"], sc["raw_coords"], [clicks], [times], dtype=torch.float64)
    m = _gpu_model(7)
    _, layers = _run_gpu(m, coords, sc["feats"], sc["raw_coords"], [clicks], [times])
    for l in range(3):
        e = rel_err(layers[l][0].cpu().numpy(), ref_layers[l][0].numpy())
        assert e < 1e-3, f"layer {l}: rel err {e}"


def test_two_batches_in_flight_match_sequential():
    """The serving pattern of bench.py: consecutive batches on alternating caller streams, sharing one model (its side
    streams, workspaces and weight caches).  Every step must reproduce, bit for bit, what the same batch gives alone."""
    from agile3d_b200.scenes import make_clicks, make_scene
    import agile3d_b200
    m = _gpu_model(7)
    batches = []
    for k, (n, seed) in enumerate([(7000, 41), (9000, 42)]):
        scs = [make_scene(n + 500 * j, 0.02, seed=seed + 10 * j, n_box=8) for j in range(3)]
        ck, tm = zip(*[make_clicks(sc, 3, 2, 1, seed=seed + j)[:2] for j, sc in enumerate(scs)])
        coords = np.concatenate([np.concatenate([np.full((sc["coords"].shape[0], 1), j, np.int32), sc["coords"]], 1)
                                 for j, sc in enumerate(scs)], 0)
        feats = np.concatenate([sc["feats"] for sc in scs], 0)
        raw = np.concatenate([sc["raw_coords"] for sc in scs], 0)
        batches.append((torch.from_numpy(coords).to(DEV), torch.from_numpy(feats).to(DEV), torch.from_numpy(raw).to(DEV),
                        list(ck), list(tm)))

    def step(b):
        c, f, r, ck, tm = batches[b]
        x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=DEV)
        out = m.forward_mask(*m.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)
        return [p.clone() for p in out["pred_masks"]] + [p.clone() for a in out["aux_outputs"] for p in a["pred_masks"]]

    ref = [step(0), step(1)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for i in range(8):
        with torch.cuda.stream(streams[i % 2]):
            outs.append((i % 2, step(i % 2)))
    torch.cuda.synchronize()
    for b, got in outs:
        for a, e in zip(got, ref[b]):
            assert torch.equal(a, e)


def test_forward_mask_is_repeatable_and_handles_unmutated():
    g = load_golden("g3000_k3")
    m = _gpu_model(g["wseed"])
    h, layers = _run_gpu(m, g["coords"], g["feats"], g["raw_coords"], [g["clicks"]], [g["times"]])
    before = h[0].F.clone()
    out2 = m.forward_mask(*h, [g["clicks"]], [g["times"]])
    assert torch.equal(h[0].F, before)
    assert torch.equal(out2["pred_masks"][0], layers[2][0])


# ------------------------------------------------------------------------------------------------ full size
def test_full_size_properties_150k():
    """BASELINE.json headline shape (150k voxels, 10 clicks -> 20 queries): properties that need no CPU oracle."""
    from agile3d_b200.backbone import CoordinateMaps
    from agile3d_b200.scenes import make_clicks, make_scene
    sc = make_scene(150000, 0.02, seed=1000 * 2 + 0)
    n = sc["coords"].shape[0]
    assert abs(n - 150000) <= 3000
    coords = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), sc["coords"]], 1)).to(DEV)
    maps = CoordinateMaps(coords, count_pairs=True)
    nbr = maps.k3[0]
    # centre offset is the identity; the 3x3x3 map is symmetric: nbr[k][o] = i  <=>  nbr[26-k][i] = o
    assert torch.equal(nbr[13], torch.arange(n, dtype=torch.int32, device=DEV))
    for k in (0, 5, 12):
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        i = nbr[k][o].long()
        assert torch.equal(nbr[26 - k][i], o.to(torch.int32))
    # every fine voxel has exactly one parent entry in the transposed map; parents partition the level
    assert int((maps.up[0] >= 0).sum()) == n
    assert int((maps.down[0] >= 0).sum()) == n
    assert maps.sizes[0] > maps.sizes[1] > maps.sizes[2] > maps.sizes[3] > maps.sizes[4] > 0
    # determinism of the canonical row order
    maps2 = CoordinateMaps(coords)
    assert all(torch.equal(a, b) for a, b in zip(maps.coords, maps2.coords))
    assert torch.equal(maps.k3[1], maps2.k3[1])
    # whole model: finite, repeatable, logits columns = 1 + K
    clicks, times, _ = make_clicks(sc, 5, 2, 0, seed=0)
    m = _gpu_model(5)
    h, layers = _run_gpu(m, coords.cpu().numpy(), sc["feats"], sc["raw_coords"], [clicks], [times])
    assert layers[2][0].shape == (n, 6) and bool(torch.isfinite(layers[2][0]).all())
    _, layers2 = _run_gpu(m, coords.cpu().numpy(), sc["feats"], sc["raw_coords"], [clicks], [times])
    assert torch.equal(layers2[2][0], layers[2][0])
    # clicked voxels of object j carry the object's own click feature: the scene is consistent row-wise
    assert h[0].F.shape == (n, 128)


# ------------------------------------------------------------------------------------------------ BASELINE shapes
# BASELINE.json configs at their FULL sizes against the CPU oracle (the largest oracle-compared scene
# was 9 000 voxels, so the split-K plans, multi-tile CTAs and 4 000+-CTA grids of the real shapes were unpinned).
#   headline : 150k voxels @2cm, 5 objects x 2 clicks (Nq = 20)           SURVEY.md §8(d) "headline" / configs[2] scene
#   c2       : 150k voxels @2cm, 1 object, 3 fg + 2 bg clicks (Nq = 15)   configs[1]
#   c5_nq30  : 80k voxels @5cm outdoor, 10 objects x 2 clicks (Nq = 30)   configs[4], the 20-click point
#   c5_nq210 : same scene, 10 objects x 20 clicks (Nq = 210)              configs[4], the end of the click loop
BASELINE_SHAPES = {
    "headline": dict(target=150000, voxel=0.02, seed=2000, outdoor=False, k=5, cpo=2, bg=0),
    "c2": dict(target=150000, voxel=0.02, seed=2001, outdoor=False, k=1, cpo=3, bg=2),
    "c5_nq30": dict(target=80000, voxel=0.05, seed=5000, outdoor=True, k=10, cpo=2, bg=0),
    "c5_nq210": dict(target=80000, voxel=0.05, seed=5000, outdoor=True, k=10, cpo=20, bg=0),
}


def _baseline_scene(cfg):
    from agile3d_b200.scenes import make_clicks, make_scene
    sc = make_scene(cfg["target"], cfg["voxel"], seed=cfg["seed"], outdoor=cfg["outdoor"])
    clicks, times, _ = make_clicks(sc, cfg["k"], cfg["cpo"], cfg["bg"], seed=cfg["seed"])
    coords = np.concatenate([np.zeros((sc["coords"].shape[0], 1), np.int32), sc["coords"]], 1)
    return sc, coords, clicks, times


@pytest.mark.parametrize("shape", ["headline", "c5_nq30"])
def test_baseline_shape_maps_bit_exact(shape):
    """Coordinate levels, parents and every kernel map of the U-Net at the full BASELINE size, bit-exact."""
    from agile3d_b200.backbone import CoordinateMaps
    _, coords, _, _ = _baseline_scene(BASELINE_SHAPES[shape])
    maps = CoordinateMaps(torch.from_numpy(coords).to(DEV))
    levels, parents = [torch.from_numpy(coords)], []
    for lvl in range(4):
        c, _, _, par = emulate.downsample(levels[-1], 2 << lvl)
        levels.append(c)
        parents.append(par)
    for lvl in range(5):
        assert torch.equal(maps.coords[lvl].cpu(), levels[lvl]), f"coords level {lvl}"
        assert torch.equal(maps.k3[lvl].cpu(), emulate.kernel_map(levels[lvl], levels[lvl], 0, 3, 1 << lvl)), f"k3 {lvl}"
    for lvl in range(4):
        assert torch.equal(maps.parents[lvl].cpu(), parents[lvl]), f"parents level {lvl}"
        assert torch.equal(maps.down[lvl].cpu(), emulate.kernel_map(levels[lvl + 1], levels[lvl], 0, 2, 1 << lvl))
        assert torch.equal(maps.up[lvl].cpu(), emulate.kernel_map_transposed(levels[lvl], parents[lvl], 1 << lvl))


@pytest.mark.parametrize("shape", list(BASELINE_SHAPES))
def test_baseline_shape_logits_vs_fp64_oracle(shape):
    """Backbone features and the mask logits of all three decoder layers against the fp64 CPU oracle at the full
    BASELINE sizes, default (tensor-core) mode, 1e-3 (max|a-b| / max|b|).  The decoder takes a discrete decision per
    voxel between layers; the comparison of layers 1 and 2 is made against the oracle run with the SAME decisions
    (helpers.decision_forced_errors), and every differing decision must be a genuine near-tie of the oracle."""
    cfg = BASELINE_SHAPES[shape]
    sc, coords, clicks, times = _baseline_scene(cfg)
    nq = 10 + cfg["k"] * cfg["cpo"] + cfg["bg"]
    m = _gpu_model(5)
    h, layers = _run_gpu(m, coords, sc["feats"], sc["raw_coords"], [clicks], [times])
    assert layers[2][0].shape == (coords.shape[0], 1 + cfg["k"])
    r = decision_forced_errors(oracle_model(5, torch.float64), coords, sc["feats"], sc["raw_coords"], [clicks], [times], layers)
    e = rel_err(h[0].F.cpu().numpy(), r["pcd"].F.numpy())
    print(f"{shape}: Nv={coords.shape[0]} Nq={nq} backbone {e:.2e}; logits vs oracle with the same decisions "
          f"{['%.2e' % v for v in r['forced']]}, vs free-running oracle {['%.2e' % v for v in r['free']]}, "
          f"differing decisions {r['flips']} (not near-ties: {r['bad_flips']})")
    assert e < 1e-3, f"{shape}: backbone features rel err {e}"
    assert r["free"][0] < 1e-3 and max(r["forced"]) < 1e-3, (shape, r["forced"], r["free"])
    assert r["bad_flips"] == 0 and max(r["flips"]) < 1e-3 * coords.shape[0], (shape, r["flips"], r["bad_flips"])
