"""Diagnostics for the tensor-core s2c kernel against the fp32 SIMT kernel and the fp64 contract emulation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import emulate  # noqa: E402
from agile3d_b200 import ops  # noqa: E402

DEV = "cuda"


def inputs(nv, nq, n_obj, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((nv, 128), generator=g)
    pos = torch.randn((nv, 128), generator=g) * 0.7
    A = torch.randn((8 * nq, 128), generator=g) * 0.05
    c = torch.randn(8 * nq, generator=g) * 0.1
    U = torch.randn((8 * nq, 128), generator=g) * 0.3
    bo, lw, lb = torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    E = torch.randn((nq, 128), generator=g) * 0.2
    n_fg = nq - 10
    q_obj = torch.tensor(sorted((i % (n_obj - 1)) + 1 for i in range(n_fg)) + [0] * 10, dtype=torch.int32)
    return x, pos, A, c, U, bo, lw, lb, E, q_obj


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def case(nv, nq, n_obj, seed=0):
    x, pos, A, c, U, bo, lw, lb, E, q_obj = inputs(nv, nq, n_obj, seed)
    d = lambda t: t.double()
    ry, rl, rlab, rcnt = emulate.s2c_mask_fwd(d(x), d(pos), d(A), d(c), d(U), d(bo), d(lw), d(lb), 1e-5, d(E), q_obj, nq, 8, n_obj)
    t = lambda v: v.to(DEV)
    args = (t(x), t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8, n_obj)
    y, lg, lab, cnt = ops.s2c_mask_fwd(*args, algo=ops.ALGO_TC)
    torch.cuda.synchronize()
    ey, el = rel(y.cpu().double(), ry), rel(lg.cpu().double(), rl)
    top2 = torch.topk(rl, 2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 1e-3
    lab_ok = bool(torch.equal(lab.cpu()[safe], rlab[safe]))
    ok = ey < 1e-4 and el < 1e-4 and lab_ok and int(cnt.sum()) == nv
    print(f"[{'OK ' if ok else 'BAD'}] nv={nv} nq={nq} n_obj={n_obj}: y rel {ey:.2e} logits rel {el:.2e} labels {lab_ok} count {int(cnt.sum())}", flush=True)
    if not ok:
        e = (y.cpu().double() - ry).abs().numpy()
        bad_rows = np.nonzero(e.max(1) > 1e-3)[0]
        bad_cols = np.nonzero(e.max(0) > 1e-3)[0]
        print(f"      y bad rows {len(bad_rows)}/{nv} first {bad_rows[:10].tolist()}; bad cols {len(bad_cols)}/128 first {bad_cols[:10].tolist()}")
        print("      y[0,:6]  ", np.round(y[0, :6].cpu().numpy(), 4).tolist(), " ref ", np.round(ry[0, :6].numpy(), 4).tolist())
        print("      lg[0]    ", np.round(lg[0].cpu().numpy(), 4).tolist(), " ref ", np.round(rl[0].numpy(), 4).tolist())
    return ok


def timing(nv, nq, n_obj):
    x, pos, A, c, U, bo, lw, lb, E, q_obj = inputs(nv, nq, n_obj, 1)
    t = lambda v: v.to(DEV)
    args = (t(x), t(pos), t(A), t(c), t(U), t(bo), t(lw), t(lb), 1e-5, t(E), t(q_obj), nq, 8, n_obj)
    for algo, name in ((ops.ALGO_SIMT, "simt"), (ops.ALGO_TC, "tc")):
        for _ in range(2):
            ops.s2c_mask_fwd(*args, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.s2c_mask_fwd(*args, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nbytes = 4 * nv * 128 * 4 + 4 * nv * n_obj + nq * nv
        print(f"timing s2c nv={nv} nq={nq}: {name} {ms:.3f} ms ({nbytes/ms/1e6:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    oks = [case(128, 16, 4), case(128, 11, 2), case(300, 20, 6), case(5003, 15, 3), case(20000, 20, 6),
           case(4100, 25, 9), case(777, 32, 12), case(150000, 20, 6)]
    print("ALL OK" if all(oks) else "SOME BAD")
    timing(150000, 20, 6)
    timing(150000, 15, 2)
