"""CUDA-event and host wall time of the two phases of one eval step (forward_backbone, forward_mask) at the bench batch
size, next to the per-family kernel sums: where is the GPU idle?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200 import ops  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate, make_inputs  # noqa: E402

dev = torch.device("cuda", 0)
model = agile3d_b200.build_model(default_args()).eval()
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev)
c, f, r, ck, tm = collate(make_inputs(8, 2000))
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step(record=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    ev[0].record()
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    t1 = time.perf_counter()
    ev[1].record()
    h = model.forward_backbone(x, raw_coordinates=r)
    t2 = time.perf_counter()
    ev[2].record()
    out = model.forward_mask(*h, click_idx=ck, click_time_idx=tm)
    t3 = time.perf_counter()
    ev[3].record()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    if record is not None:
        record.append(([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], [1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))]))
    return out


for _ in range(3):
    step()
rec = []
for _ in range(5):
    step(rec)
g = [sum(x[0][i] for x in rec) / len(rec) for i in range(3)]
h = [sum(x[1][i] for x in rec) / len(rec) for i in range(4)]
print(f"GPU ms   : SparseTensor {g[0]:.2f}  forward_backbone {g[1]:.2f}  forward_mask {g[2]:.2f}  total {sum(g):.2f}")
print(f"host ms  : SparseTensor {h[0]:.2f}  forward_backbone {h[1]:.2f}  forward_mask {h[2]:.2f}  final sync {h[3]:.2f}")
prof = ops.Profiler()
ops.set_profiler(prof)
step()
ops.set_profiler(None)
fam = prof.summary()
print("kernel families (ms):", {k: round(v["ms"], 2) for k, v in fam.items()})
