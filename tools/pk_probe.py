"""One full-size sparse conv (8 x 150k voxels, level-0 3x3x3 map) through the pair-packed kernel under its experiment
switches (AG3D_PK_NA / AG3D_PK_NB / AG3D_PK_DEBUG, read once per process).  Usage: pk_probe.py cin cout [dense]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.backbone import CoordinateMaps  # noqa: E402
from agile3d_b200.scenes import make_scene  # noqa: E402

cin, cout = int(sys.argv[1]), int(sys.argv[2])
algo = ops.ALGO_TC if len(sys.argv) > 3 else ops.ALGO_TC_PACKED
scs = [make_scene(150000, 0.02, seed=2000 + b) for b in range(8)]
coords = np.concatenate([np.concatenate([np.full((s["coords"].shape[0], 1), b, np.int32), s["coords"]], 1)
                         for b, s in enumerate(scs)], 0)
maps = CoordinateMaps(torch.from_numpy(coords).cuda())
nbr, n = maps.k3[0], maps.sizes[0]
g = torch.Generator().manual_seed(1)
x = ops.pack_split(torch.randn((n, cin), generator=g).cuda())
w = (torch.randn((27, cin, cout), generator=g) * 0.05).cuda()
wtc = ops.prepare_tc_weight(w)
out = torch.zeros((n, cout), device="cuda")
run = lambda: ops.spconv_fwd(x, nbr, w, out, relu=True, algo=algo, weight_tc=wtc, in_split=True, out_split=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{cin}->{cout} {'dense' if len(sys.argv) > 3 else 'packed'} NA={os.environ.get('AG3D_PK_NA', '-')} NB={os.environ.get('AG3D_PK_NB', '-')} "
      f"DEBUG={os.environ.get('AG3D_PK_DEBUG', '0')}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
