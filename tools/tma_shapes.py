"""Tensor-core sparse conv (split rows, default gather engine) against the exact fp32 kernel over the backbone's layer
shapes and row counts; stops at the first CUDA error and says which shape raised it.  Usage: tma_shapes.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402

SHAPES = [(32, 32, 27), (32, 64, 27), (64, 64, 27), (64, 128, 27), (128, 128, 27), (128, 256, 27), (256, 256, 27),
          (384, 256, 27), (192, 128, 27), (128, 96, 27), (96, 96, 27), (32, 32, 8), (64, 64, 8), (256, 256, 8),
          (128, 96, 1), (96, 128, 1), (384, 256, 1)]
ns = [int(a) for a in sys.argv[1:]] or [6, 50, 122, 850, 3600, 15000]
g = torch.Generator().manual_seed(1)
for n in ns:
    for cin, cout, K in SHAPES:
        n_in = max(4, n // 2 if K == 8 else n)
        x = torch.randn((n_in, cin), generator=g).cuda()
        w = (torch.randn((K, cin, cout), generator=g) * 0.05).cuda()
        if K == 1:
            nbr = None
            x = torch.randn((n, cin), generator=g).cuda()
        else:
            nbr = torch.randint(0, n_in, (K, n), generator=g, dtype=torch.int32)
            nbr[torch.rand((K, n), generator=g) > 0.46] = -1
            nbr = nbr.cuda()
        ref = torch.empty((n, cout), device="cuda")
        ops.spconv_fwd(x, nbr, w, ref, relu=True, algo=ops.ALGO_SIMT)
        out = torch.empty((n, cout), device="cuda")
        tag = f"n={n} cin={cin} cout={cout} K={K}"
        try:
            ops.spconv_fwd(ops.pack_split(x), nbr, w, out, relu=True, algo=ops.ALGO_TC, weight_tc=ops.prepare_tc_weight(w),
                           in_split=True, out_split=True)
            torch.cuda.synchronize()
            got = ops.unpack_split(out)
            err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
            print(f"{tag}: rel err {err:.2e}" + ("" if err < 2e-4 else "   <-- WRONG"), flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{tag}: FAILED {str(e)[:120]}", flush=True)
            sys.exit(1)
print("all shapes ok")
