"""Pair-packed tensor-core sparse conv (csrc/spconv_pk.cu) against the exact fp32 kernel and, for time, against the
dense-tile tensor-core kernel, over the backbone's large-level layer shapes.  Neighbour tables: a real 3x3x3 / 2x2x2
map of a synthetic scene (default) or random.  Usage: pk_shapes.py [voxels ...]   (AG3D_PK_REPS=n timing repeats)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.backbone import CoordinateMaps  # noqa: E402
from agile3d_b200.scenes import make_scene  # noqa: E402

SHAPES = [(96, 96, 27), (128, 96, 27), (32, 32, 27), (64, 64, 27), (128, 128, 27), (192, 128, 27), (96, 96, "up"),
          (32, 32, "down"), (128, 96, "up")]
reps = int(os.environ.get("AG3D_PK_REPS", "10"))
ns = [int(a) for a in sys.argv[1:]] or [700, 20000, 150000]
g = torch.Generator().manual_seed(1)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in ns:
    batch = 8 if n >= 100000 else 1
    scs = [make_scene(n, 0.02, seed=2000 + b) for b in range(batch)]
    coords = np.concatenate([np.concatenate([np.full((s["coords"].shape[0], 1), b, np.int32), s["coords"]], 1)
                             for b, s in enumerate(scs)], 0)
    maps = CoordinateMaps(torch.from_numpy(coords).cuda())
    for cin, cout, kind in SHAPES:
        if kind == 27:
            nbr, n_in, n_out = maps.k3[0], maps.sizes[0], maps.sizes[0]
        elif kind == "up":
            nbr, n_in, n_out = maps.up[0], maps.sizes[1], maps.sizes[0]
        else:
            nbr, n_in, n_out = maps.down[0], maps.sizes[0], maps.sizes[1]
        K = nbr.shape[0]
        x = torch.randn((n_in, cin), generator=g).cuda()
        w = (torch.randn((K, cin, cout), generator=g) * 0.05).cuda()
        res = torch.randn((n_out, cout), generator=g).cuda()
        sc, sh = (torch.rand(cout, generator=g) + 0.5).cuda(), torch.randn(cout, generator=g).cuda()
        ref = torch.empty((n_out, cout), device="cuda")
        ops.spconv_fwd(x, nbr, w, ref, sc, sh, residual=res, relu=True, algo=ops.ALGO_SIMT)
        xs, wtc, rs = ops.pack_split(x), ops.prepare_tc_weight(w), ops.pack_split(res)
        tag = f"rows {n_out} (in {n_in}) cin={cin} cout={cout} K={K}"
        outs = {}
        try:
            for name, algo in (("packed", ops.ALGO_TC_PACKED), ("dense", ops.ALGO_TC)):
                out = torch.zeros((n_out, cout), device="cuda")
                run = lambda: ops.spconv_fwd(xs, nbr, w, out, sc, sh, residual=rs, relu=True, algo=algo, weight_tc=wtc,
                                             in_split=True, out_split=True, res_split=True)
                ms = timed(run)
                got = ops.unpack_split(out)
                outs[name] = (ms, float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
            # fp32 output / fp32 residual variant of the packed kernel
            out = torch.zeros((n_out, cout), device="cuda")
            ops.spconv_fwd(xs, nbr, w, out, None, None, residual=res, relu=False, algo=ops.ALGO_TC_PACKED, weight_tc=wtc, in_split=True)
            ref2 = torch.empty((n_out, cout), device="cuda")
            ops.spconv_fwd(x, nbr, w, ref2, None, None, residual=res, relu=False, algo=ops.ALGO_SIMT)
            e2 = float((out - ref2).abs().max() / ref2.abs().max().clamp_min(1e-30))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"{tag}: FAILED {str(e)[:200]}", flush=True)
            sys.exit(1)
        bad = "" if max(outs["packed"][1], e2) < 2e-4 else "   <-- WRONG"
        print(f"{tag}: packed {outs['packed'][0]:.4f} ms err {outs['packed'][1]:.2e} (fp32 out {e2:.2e}) | dense "
              f"{outs['dense'][0]:.4f} ms err {outs['dense'][1]:.2e} | speed-up {outs['dense'][0] / outs['packed'][0]:.2f}{bad}", flush=True)
print("done")
