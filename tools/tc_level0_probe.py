"""One real level-0 sparse-conv layer (8 x 150k voxels, internal row order, split rows) under the kernel's experiment switches
(AG3D_TC_DEBUG: 1 no MMAs, 2 no gathers, 64 no epilogue).  Usage: AG3D_TC_DEBUG=.. python tools/tc_level0_probe.py [cin cout]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.backbone import CoordinateMaps  # noqa: E402
from bench import collate, make_inputs  # noqa: E402

cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 96)
dev = "cuda"
c, f, r, ck, tm = collate(make_inputs(8, 2000, 150000))
maps = CoordinateMaps(c.to(dev), reorder=True)
n = maps.sizes[0]
g = torch.Generator().manual_seed(1)
x = ops.pack_split_rows(torch.randn((n, cin), generator=g).to(dev))
w = (torch.randn((27, cin, cout), generator=g) * 0.03).to(dev)
wtc = ops.prepare_tc_weight(w)
out = torch.empty((n, cout), device=dev)
sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)


def run():
    ops.spconv_fwd(x, maps.k3[0], w, out, sc, sh, relu=True, algo=ops.ALGO_TC, weight_tc=wtc, in_split=True, out_split=True)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"level 0, {n} rows, {cin}->{cout}, K=27, AG3D_TC_DEBUG={os.environ.get('AG3D_TC_DEBUG', '0')}: {e0.elapsed_time(e1) / 10:.4f} ms")
