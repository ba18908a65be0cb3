"""Diagnostics for the tensor-core c2s kernel against the fp64 contract emulation (and the fp32 SIMT kernel)."""
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
    qf = torch.randn((8 * nq, 128), generator=g) * 0.08
    n_fg = nq - 10
    q_obj = torch.tensor(sorted((i % (n_obj - 1)) + 1 for i in range(n_fg)) + [0] * 10, dtype=torch.int32)
    label = torch.randint(0, max(n_obj - 1, 1), (nv,), generator=g).to(torch.uint8)
    cnt = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
    return x, pos, qf, q_obj, label, cnt


def report(name, got, ref):
    err = (got - ref).abs()
    rel = float(err.max() / ref.abs().max())
    ok = rel < 1e-4
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel {rel:.2e}", flush=True)
    if not ok:
        e = err.numpy()
        bad_rows = np.nonzero(e.max(1) > 1e-3 * float(ref.abs().max()))[0]
        bad_cols = np.nonzero(e.max(0) > 1e-3 * float(ref.abs().max()))[0]
        print(f"      bad rows {len(bad_rows)}/{e.shape[0]} first {bad_rows[:12].tolist()}; bad cols {len(bad_cols)}/128 first {bad_cols[:12].tolist()}")
        print("      got[0,:6]", np.round(got[0, :6].numpy(), 4).tolist(), "ref", np.round(ref[0, :6].numpy(), 4).tolist())
        with np.errstate(divide="ignore", invalid="ignore"):
            print("      median got/ref", float(np.nanmedian((got / ref).numpy())))
    return ok


def case(nv, nq, n_obj, seed=0):
    x, pos, qf, q_obj, label, cnt = inputs(nv, nq, n_obj, seed)
    t = lambda v: v.to(DEV)
    ref = emulate.c2s_attn_fwd(x.double(), pos.double(), qf.double(), nq, 8)
    got = ops.c2s_attn_fwd(t(x), t(pos), t(qf), nq, 8, algo=ops.ALGO_TC).cpu().double()
    a = report(f"nv={nv} nq={nq} unmasked", got, ref)
    ref = emulate.c2s_attn_fwd(x.double(), pos.double(), qf.double(), nq, 8, label, q_obj, cnt)
    got = ops.c2s_attn_fwd(t(x), t(pos), t(qf), nq, 8, t(label), t(q_obj), t(cnt), algo=ops.ALGO_TC).cpu().double()
    b = report(f"nv={nv} nq={nq} masked n_obj={n_obj}", got, ref)
    return a and b


def timing(nv, nq, n_obj):
    x, pos, qf, q_obj, label, cnt = inputs(nv, nq, n_obj, 1)
    t = lambda v: v.to(DEV)
    args = (t(x), t(pos), t(qf), nq, 8, t(label), t(q_obj), t(cnt))
    for algo, name in ((ops.ALGO_SIMT, "simt"), (ops.ALGO_TC, "tc")):
        for _ in range(2):
            ops.c2s_attn_fwd(*args, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.c2s_attn_fwd(*args, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"timing c2s nv={nv} nq={nq}: {name} {ms:.3f} ms ({(4*nv*256 + nq*nv)/ms/1e6:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    oks = [case(64, 16, 3), case(64, 11, 2), case(200, 12, 2), case(5003, 15, 3), case(20000, 20, 6), case(4100, 27, 9),
           case(3000, 45, 11), case(150000, 20, 6)]
    print("ALL OK" if all(oks) else "SOME BAD")
    timing(150000, 20, 6)
    timing(150000, 15, 2)
