"""CUDA-event times of the decoder's two voxel-streaming kernels, fp32-row variants vs split-row (TMA-fed) variants, on one
150k-voxel scene with 20 click queries.  Usage: python tools/dec_split_time.py [--voxels N] [--nq Q]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxels", type=int, default=150000)
ap.add_argument("--nq", type=int, default=20)
a = ap.parse_args()
dev = "cuda"
g = torch.Generator().manual_seed(0)
nv, nq, n_obj = a.voxels, a.nq, 6
x = torch.randn((nv, 128), generator=g).to(dev)
pos = (torch.randn((nv, 128), generator=g) * 0.7).to(dev)
qf = (torch.randn((8 * nq, 128), generator=g) * 0.08).to(dev)
A = (torch.randn((8 * nq, 128), generator=g) * 0.05).to(dev)
c = (torch.randn(8 * nq, generator=g) * 0.1).to(dev)
U = (torch.randn((8 * nq, 128), generator=g) * 0.3).to(dev)
bo, lw, lb = (torch.randn(128, generator=g) * 0.1).to(dev), (torch.rand(128, generator=g) + 0.5).to(dev), (torch.randn(128, generator=g) * 0.1).to(dev)
E = (torch.randn((nq, 128), generator=g) * 0.2).to(dev)
q_obj = torch.tensor(sorted((i % (n_obj - 1)) + 1 for i in range(nq - 10)) + [0] * 10, dtype=torch.int32, device=dev)
label = torch.randint(0, n_obj, (nv,), generator=g).to(torch.uint8).to(dev)
cnt = torch.bincount(label.long(), minlength=n_obj).to(torch.int32)
xs, ps = ops.pack_split_rows(x), ops.pack_split_rows(pos)
flush = torch.empty(2 << 30, dtype=torch.uint8, device=dev)     # ~0.7 ms of memset: the host enqueues fn() behind it


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


gb = lambda b, ms: b / ms / 1e6
c2b, s2b = 4 * nv * 128 * 2, 4 * nv * 128 * 4
if os.environ.get("ONLY_C2S"):
    ms = timed(lambda: ops.c2s_attn_fwd(xs, ps, qf, nq, 8, label, q_obj, cnt, split=True))
    print(f"c2s split debug={os.environ.get('AG3D_C2S_DEBUG', '0')}: {ms:.4f} ms")
    sys.exit(0)
if os.environ.get("ONLY_S2C"):
    ms = timed(lambda: ops.s2c_mask_fwd(xs, ps, A, c, U, bo, lw, lb, 1e-5, E, q_obj, nq, 8, n_obj, split=True))
    print(f"s2c split debug={os.environ.get('AG3D_S2C_DEBUG', '0')}: {ms:.4f} ms")
    sys.exit(0)
for name, fn, b in [
    ("c2s fp32 rows        ", lambda: ops.c2s_attn_fwd(x, pos, qf, nq, 8, label, q_obj, cnt, algo=2), c2b),
    ("c2s split rows (TMA) ", lambda: ops.c2s_attn_fwd(xs, ps, qf, nq, 8, label, q_obj, cnt, split=True), c2b),
    ("s2c fp32 rows        ", lambda: ops.s2c_mask_fwd(x, pos, A, c, U, bo, lw, lb, 1e-5, E, q_obj, nq, 8, n_obj, algo=2), s2b),
    ("s2c split rows (TMA) ", lambda: ops.s2c_mask_fwd(xs, ps, A, c, U, bo, lw, lb, 1e-5, E, q_obj, nq, 8, n_obj, split=True), s2b),
    ("s2c split, no x write", lambda: ops.s2c_mask_fwd(xs, ps, A, c, U, bo, lw, lb, 1e-5, E, q_obj, nq, 8, n_obj, split=True, write_x=False), s2b),
]:
    ms = timed(fn)
    print(f"{name}: {ms:.4f} ms  {gb(b, ms):7.1f} GB/s algorithmic ({nv} voxels, {nq} queries; includes the merge / prep launch)")
