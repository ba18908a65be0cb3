"""Diagnostic ladder for the tcgen05 sparse-conv path: every case is compared with the exact-fp32 SIMT kernel on the
same inputs; prints error statistics and, on mismatch, where the error sits (rows / columns / tiles)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402

DEV = "cuda"


def rand_map(n_out, n_in, K, density, g):
    nbr = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
    drop = torch.rand((K, n_out), generator=g) > density
    nbr[drop] = -1
    return nbr


def case(name, n_out, n_in, K, cin, cout, density=0.5, residual=False, relu=False, slices=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n_in, cin), generator=g)
    w = torch.randn((K, cin, cout), generator=g) / np.sqrt(cin * max(1.0, K * density))
    nbr = None if (K == 1 and n_in == n_out and density >= 1.0) else rand_map(n_out, n_in, K, density, g)
    scale = (torch.rand(cout, generator=g) + 0.5) if relu else None
    shift = torch.randn(cout, generator=g) * 0.1 if relu else None
    res = torch.randn((n_out, cout), generator=g) if residual else None
    xd = x.to(DEV)
    if slices:
        buf = torch.zeros((n_in, cin + 32), device=DEV)
        buf[:, 32:] = xd
        xd = buf[:, 32:]
    wd = w.to(DEV)
    wtc = ops.prepare_tc_weight(wd)
    nd = nbr.to(DEV) if nbr is not None else None
    sd, hd, rd = (t.to(DEV) if t is not None else None for t in (scale, shift, res))
    ref = torch.empty((n_out, cout), device=DEV)
    ops.spconv_fwd(xd, nd, wd, ref, sd, hd, rd, relu=relu, algo=ops.ALGO_SIMT)
    outb = torch.full((n_out, cout + (64 if slices else 0)), -7.0, device=DEV)
    out = outb[:, 64:] if slices else outb
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops.spconv_fwd(xd, nd, wd, out, sd, hd, rd, relu=relu, algo=ops.ALGO_TC, weight_tc=wtc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = (out - ref).abs()
    scale_ref = float(ref.abs().max())
    rel = float(err.max()) / max(scale_ref, 1e-30)
    ok = rel < 2e-4 and (not slices or bool((outb[:, :64] == -7.0).all()))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: n_out={n_out} K={K} {cin}->{cout} rel_err={rel:.3e} "
          f"max|ref|={scale_ref:.3f} first-call {dt*1e3:.2f} ms", flush=True)
    if not ok:
        e = err.cpu().numpy()
        bad_rows = np.nonzero(e.max(1) > 1e-3 * scale_ref)[0]
        bad_cols = np.nonzero(e.max(0) > 1e-3 * scale_ref)[0]
        print(f"      bad rows: {len(bad_rows)}/{n_out} first {bad_rows[:12].tolist()} ; bad cols: {len(bad_cols)}/{cout} "
              f"first {bad_cols[:12].tolist()}")
        o, r = out.cpu().numpy(), ref.cpu().numpy()
        print("      out[0,:8] ", np.round(o[0, :8], 4).tolist())
        print("      ref[0,:8] ", np.round(r[0, :8], 4).tolist())
        rr = bad_rows[0] if len(bad_rows) else 0
        print(f"      out[{rr},:8]", np.round(o[rr, :8], 4).tolist())
        print(f"      ref[{rr},:8]", np.round(r[rr, :8], 4).tolist())
        # does the output match the reference of some permutation? ratio statistics
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = o / r
        print("      median out/ref ratio:", float(np.nanmedian(ratio)))
    return ok


def main():
    oks = []
    oks.append(case("identity 1 tile", 128, 128, 1, 32, 32, density=1.0))
    oks.append(case("identity n=96", 128, 128, 1, 32, 96, density=1.0))
    oks.append(case("identity 3 slabs, 3 tiles, tail", 300, 300, 1, 96, 96, density=1.0))
    oks.append(case("identity 256 cols", 700, 700, 1, 128, 256, density=1.0))
    oks.append(case("K=27 sparse map", 1000, 1000, 27, 64, 64, density=0.45, seed=1))
    oks.append(case("K=27 epilogue + slices", 5000, 5000, 27, 128, 256, density=0.45, residual=True, relu=True,
                    slices=True, seed=2))
    oks.append(case("K=8 down", 3000, 11000, 8, 32, 32, density=0.5, relu=True, seed=3))
    oks.append(case("K=8 wide", 9000, 2500, 8, 256, 128, density=0.125, relu=True, seed=4))
    oks.append(case("K=27 very sparse (offset skipping)", 2000, 2000, 27, 96, 96, density=0.002, seed=5))
    oks.append(case("384->256", 3000, 3000, 27, 384, 256, density=0.4, residual=True, relu=True, seed=6))
    oks.append(case("split-K: 400 rows 256->256", 400, 400, 27, 256, 256, density=0.4, residual=True, relu=True, seed=11))
    oks.append(case("split-K: 2300 rows 384->256", 2300, 2300, 27, 384, 256, density=0.4, residual=True, relu=True,
                    slices=True, seed=12))
    oks.append(case("split-K: K=8 up 400->1800 rows", 1800, 400, 8, 256, 256, density=0.125, relu=True, seed=13))
    oks.append(case("full size 96->96", 150000, 150000, 27, 96, 96, density=0.46, residual=True, relu=True, seed=7))
    print("ALL OK" if all(oks) else "SOME BAD")
    # timing of the big case, both paths
    g = torch.Generator().manual_seed(9)
    n = 150000
    x = torch.randn((n, 96), generator=g).to(DEV)
    w = (torch.randn((27, 96, 96), generator=g) * 0.03).to(DEV)
    nbr = rand_map(n, n, 27, 0.46, g).to(DEV)
    wtc = ops.prepare_tc_weight(w)
    out = torch.empty((n, 96), device=DEV)
    for algo, name in ((ops.ALGO_SIMT, "simt"), (ops.ALGO_TC, "tc")):
        for _ in range(2):
            ops.spconv_fwd(x, nbr, w, out, algo=algo, weight_tc=wtc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.spconv_fwd(x, nbr, w, out, algo=algo, weight_tc=wtc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        pairs = int((nbr >= 0).sum())
        print(f"timing 150k 96->96 K=27 random map: {name} {ms:.3f} ms  ({2*pairs*96*96/ms/1e9:.2f} TFLOP/s algorithmic)")


if __name__ == "__main__":
    main()
