"""Tensor-core weight gradient (ag3d_spconv_bwd_weight_tc) against the fp32 SIMT kernel over the backbone's layer
shapes; prints relative errors and times.  Usage: wgrad_shapes.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402

SHAPES = [(32, 32, 27), (32, 64, 27), (64, 64, 27), (128, 128, 27), (256, 256, 27), (384, 256, 27), (192, 128, 27),
          (128, 96, 27), (96, 96, 27), (32, 32, 8), (256, 128, 8), (128, 96, 1), (96, 128, 1)]
ns = [int(a) for a in sys.argv[1:]] or [50, 1000, 20000]
g = torch.Generator().manual_seed(1)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3


for n in ns:
    for cin, cout, K in SHAPES:
        n_in = n if K != 8 else max(4, n // 3)
        x = torch.randn((n_in, cin), generator=g).cuda()
        dy = torch.randn((n, cout), generator=g).cuda()
        if K == 1:
            nbr = None
        else:
            nbr = torch.randint(0, n_in, (K, n), generator=g, dtype=torch.int32)
            nbr[torch.rand((K, n), generator=g) > 0.46] = -1
            nbr = nbr.cuda()
        ref = ops.spconv_bwd_weight(x, nbr, dy, K)
        tag = f"n={n} cin={cin} cout={cout} K={K}"
        try:
            xs, ds = ops.pack_split_rows(x), ops.pack_split_rows(dy)
            got = ops.spconv_bwd_weight_tc(xs, nbr, ds, K)
            torch.cuda.synchronize()
            err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
            line = f"{tag}: rel err {err:.2e}" + ("" if err < 1e-4 else "   <-- WRONG")
            if n >= 20000:
                line += f"   simt {timed(lambda: ops.spconv_bwd_weight(x, nbr, dy, K)):.3f} ms  tc {timed(lambda: ops.spconv_bwd_weight_tc(xs, nbr, ds, K)):.3f} ms"
            print(line, flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{tag}: FAILED {str(e)[:160]}", flush=True)
            sys.exit(1)
print("done")
