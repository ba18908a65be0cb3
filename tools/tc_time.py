"""Timing + correctness of one full-size tensor-core sparse conv (150k rows, 96->96, K=27, random map), fp32 rows vs
"split" (bf16 hi/lo pair) rows.  AG3D_TC_DEBUG experiments apply (see spconv_tc.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402

DEV = "cuda"
g = torch.Generator().manual_seed(9)
n = 150000
x = torch.randn((n, 96), generator=g).to(DEV)
res = torch.randn((n, 96), generator=g).to(DEV)
w = (torch.randn((27, 96, 96), generator=g) * 0.03).to(DEV)
nbr = torch.randint(0, n, (27, n), generator=g, dtype=torch.int32)
nbr[torch.rand((27, n), generator=g) > 0.46] = -1
nbr = nbr.to(DEV)
wtc = ops.prepare_tc_weight(w)
ref = torch.empty((n, 96), device=DEV)
ops.spconv_fwd(x, nbr, w, ref, residual=res, relu=True, algo=ops.ALGO_SIMT)
xs, rs = ops.pack_split(x), ops.pack_split(res)


def run(split):
    out = torch.empty((n, 96), device=DEV)
    kw = dict(in_split=True, out_split=True, res_split=True) if split else {}
    xin, rin = (xs, rs) if split else (x, res)
    for _ in range(3):
        ops.spconv_fwd(xin, nbr, w, out, residual=rin, relu=True, algo=ops.ALGO_TC, weight_tc=wtc, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.spconv_fwd(xin, nbr, w, out, residual=rin, relu=True, algo=ops.ALGO_TC, weight_tc=wtc, **kw)
    e1.record()
    torch.cuda.synchronize()
    got = ops.unpack_split(out) if split else out
    rel = float((got - ref).abs().max() / ref.abs().max())
    print(f"AG3D_TC_DEBUG={os.environ.get('AG3D_TC_DEBUG', '0')} split={int(split)}: {e0.elapsed_time(e1) / 10:.4f} ms  rel err {rel:.2e}",
          flush=True)


run(False)
run(True)
