"""tcgen05 sparse-conv path (bf16x3 split, fp32 accumulate in TMEM) against the exact-fp32 SIMT kernel and the
CPU oracle.  Tolerance: 2e-4 relative per layer (bf16x3 drops terms of relative size <= ~1e-5 per product), and the
end-to-end 1e-3 criterion on mask logits with the tensor-core backbone."""
import numpy as np
import pytest
import torch

import emulate
from helpers import GOLDEN_CASES, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from agile3d_b200._lib import lib
    lib()


def _rand_map(n_out, n_in, K, density, g):
    nbr = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
    nbr[torch.rand((K, n_out), generator=g) > density] = -1
    return nbr


@pytest.mark.parametrize("n_out,n_in,K,cin,cout,density", [
    (128, 128, 1, 32, 32, 1.0), (300, 300, 1, 96, 96, 1.0), (700, 700, 1, 128, 256, 1.0),
    (1000, 1000, 27, 64, 64, 0.45), (5000, 5000, 27, 128, 256, 0.45), (3000, 11000, 8, 32, 32, 0.5),
    (9000, 2500, 8, 256, 128, 0.125), (2000, 2000, 27, 96, 96, 0.002), (3000, 3000, 27, 384, 256, 0.4),
    (4000, 4000, 27, 192, 128, 0.4), (40000, 40000, 27, 96, 96, 0.46),
    (400, 400, 27, 256, 256, 0.4), (2300, 2300, 27, 384, 256, 0.4), (1800, 400, 8, 256, 256, 0.125), (130, 130, 27, 128, 128, 0.5),
])
def test_spconv_tc_vs_fp32(n_out, n_in, K, cin, cout, density):
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(n_out + K + cin + cout)
    x = torch.randn((n_in, cin), generator=g).to(DEV)
    w = (torch.randn((K, cin, cout), generator=g) / np.sqrt(cin * max(1.0, K * density))).to(DEV)
    nbr = None if density >= 1.0 else _rand_map(n_out, n_in, K, density, g).to(DEV)
    sc, sh = (torch.rand(cout, generator=g) + 0.5).to(DEV), (torch.randn(cout, generator=g) * 0.1).to(DEV)
    res = torch.randn((n_out, cout), generator=g).to(DEV)
    ref = torch.empty((n_out, cout), device=DEV)
    ops.spconv_fwd(x, nbr, w, ref, sc, sh, res, relu=True, algo=ops.ALGO_SIMT)
    buf = torch.full((n_out, cout + 64), -7.0, device=DEV)
    ops.spconv_fwd(x, nbr, w, buf[:, 64:], sc, sh, res, relu=True, algo=ops.ALGO_TC, weight_tc=ops.prepare_tc_weight(w))
    assert rel_err(buf[:, 64:].cpu().numpy(), ref.cpu().numpy()) < 2e-4
    assert bool((buf[:, :64] == -7.0).all())


def test_spconv_tc_vs_cpu_oracle_real_map():
    from agile3d_b200 import ops
    rng = np.random.default_rng(0)
    c = np.unique(rng.integers(0, 40, size=(6000, 3)) // np.array([1, 1, 6]), axis=0)
    coords = torch.from_numpy(np.concatenate([np.zeros((c.shape[0], 1)), c], 1).astype(np.int32))
    nbr = emulate.kernel_map(coords, coords, 0, 3, 1)
    g = torch.Generator().manual_seed(5)
    x = torch.randn((coords.shape[0], 96), generator=g)
    w = torch.randn((27, 96, 128), generator=g) * 0.05
    ref = emulate.spconv_fwd(x.double(), nbr, w.double(), torch.empty((coords.shape[0], 128), dtype=torch.float64))
    out = torch.empty((coords.shape[0], 128), device=DEV)
    wd = w.to(DEV)
    ops.spconv_fwd(x.to(DEV), nbr.to(DEV), wd, out, algo=ops.ALGO_TC, weight_tc=ops.prepare_tc_weight(wd))
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-4


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_end_to_end_tc_backbone_vs_reference_golden(name):
    import agile3d_b200
    from agile3d_b200 import ops
    from agile3d_b200.weights import default_args, synth_state_dict
    g = load_golden(name)
    m = agile3d_b200.build_model(default_args()).eval()
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=g["wseed"]))
    m = m.to(DEV)
    m.backbone.algo = ops.ALGO_TC
    x = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(g["coords"]), features=torch.from_numpy(g["feats"]), device=DEV)
    h = m.forward_backbone(x, torch.from_numpy(g["raw_coords"]).to(DEV))
    assert rel_err(h[0].F.cpu().numpy()[::4], g["pcd_features"]) < 1e-3
    out = m.forward_mask(*h, [g["clicks"]], [g["times"]])
    layers = [a["pred_masks"][0] for a in out["aux_outputs"]] + [out["pred_masks"][0]]
    for l in range(3):
        e = rel_err(layers[l].cpu().numpy(), g["logits"][l])
        assert e < 1e-3, f"layer {l}: {e}"


@pytest.mark.parametrize("n_out,n_in,K,cin,cout,density", [
    (300, 300, 1, 96, 96, 1.0), (1000, 1000, 27, 64, 64, 0.45), (5000, 5000, 27, 128, 256, 0.45),
    (9000, 2500, 8, 256, 128, 0.125), (400, 400, 27, 256, 256, 0.4), (40000, 40000, 27, 96, 96, 0.46),
])
def test_spconv_tc_split_rows_vs_fp32(n_out, n_in, K, cin, cout, density):
    """Same layer with every feature tensor stored as bf16 hi/lo pair rows (the backbone's tensor-core format)."""
    from agile3d_b200 import ops
    g = torch.Generator().manual_seed(n_out + K + cin)
    x = torch.randn((n_in, cin), generator=g).to(DEV)
    w = (torch.randn((K, cin, cout), generator=g) / np.sqrt(cin * max(1.0, K * density))).to(DEV)
    nbr = None if density >= 1.0 else _rand_map(n_out, n_in, K, density, g).to(DEV)
    sc, sh = (torch.rand(cout, generator=g) + 0.5).to(DEV), (torch.randn(cout, generator=g) * 0.1).to(DEV)
    res = torch.randn((n_out, cout), generator=g).to(DEV)
    ref = torch.empty((n_out, cout), device=DEV)
    ops.spconv_fwd(x, nbr, w, ref, sc, sh, res, relu=True, algo=ops.ALGO_SIMT)
    # pack/unpack round trip is exact to 2^-17
    assert rel_err(ops.unpack_split(ops.pack_split(x)).cpu().numpy(), x.cpu().numpy()) < 1e-5
    xin = torch.zeros((n_in, cin + 32), device=DEV)
    xin[:, 32:] = ops.pack_split(x)
    buf = torch.full((n_out, cout + 64), -7.0, device=DEV)
    ops.spconv_fwd(xin[:, 32:], nbr, w, buf[:, 64:], sc, sh, ops.pack_split(res), relu=True, algo=ops.ALGO_TC,
                   weight_tc=ops.prepare_tc_weight(w), in_split=True, out_split=True, res_split=True)
    got = ops.unpack_split(buf[:, 64:].contiguous())
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 2e-4
    assert bool((buf[:, :64] == -7.0).all())
    # split input, fp32 output (the head convolution)
    out32 = torch.empty((n_out, cout), device=DEV)
    ops.spconv_fwd(xin[:, 32:], nbr, w, out32, sc, sh, res, relu=True, algo=ops.ALGO_TC,
                   weight_tc=ops.prepare_tc_weight(w), in_split=True)
    assert rel_err(out32.cpu().numpy(), ref.cpu().numpy()) < 2e-4


def test_stem_split_output():
    from agile3d_b200 import ops
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(0, 30, size=(4000, 3)), axis=0)
    coords = torch.from_numpy(np.concatenate([np.zeros((c.shape[0], 1)), c], 1).astype(np.int32)).to(DEV)
    g = torch.Generator().manual_seed(1)
    f = torch.rand((coords.shape[0], 3), generator=g).to(DEV)
    w = (torch.randn((125, 3, 32), generator=g) * 0.1).to(DEV)
    table, cap, _ = ops.hash_build(coords)
    a = torch.empty((coords.shape[0], 32), device=DEV)
    b = torch.empty((coords.shape[0], 32), device=DEV)
    ops.stem_conv_fwd(coords, f, table, cap, 5, w, a)
    ops.stem_conv_fwd(coords, f, table, cap, 5, w, b, out_split=True)
    assert rel_err(ops.unpack_split(b).cpu().numpy(), a.cpu().numpy()) < 1e-5
