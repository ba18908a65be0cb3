"""Kernel timeline of one data-parallel training step (torch.profiler / CUPTI, rank 0): which of this library's kernels run
WHILE the NCCL all-reduce kernels of the gradient buckets run (GradBuckets.attach: the exchange of a stage starts from
inside the backbone backward).  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/ddp_overlap_trace.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import agile3d_b200  # noqa: E402
from agile3d_b200 import dist as agd  # noqa: E402
from agile3d_b200.optim import FlatAdamW, GradBuckets  # noqa: E402
from agile3d_b200.scenes import make_clicks, make_scene  # noqa: E402
from agile3d_b200.weights import default_args, synth_state_dict  # noqa: E402
from bench import collate  # noqa: E402

rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
agd.init_from_env(backend="nccl", device=dev)
margs = default_args()
model = agile3d_b200.build_model(margs)
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
model = model.to(dev).train()
criterion = agile3d_b200.build_criterion(margs)
opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=0.1)
buckets = GradBuckets(opt, n_buckets=6)
buckets.attach(model)
batch, targets = [], []
for i in range(2):
    sc = make_scene(150000, 0.02, seed=2000 + 100 * rank + i)
    clicks, times, lab = make_clicks(sc, 5, 2, 0, seed=2000 + i)
    batch.append((sc, clicks, times))
    targets.append(torch.from_numpy(lab.astype(np.int32)).to(dev))
c, f, r, ck, tm = collate(batch)
c, f, r = c.to(dev), f.to(dev), r.to(dev)


def step():
    x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
    out = model.forward_mask(*model.forward_backbone(x, raw_coordinates=r), click_idx=ck, click_time_idx=tm)
    w = agile3d_b200.cal_click_loss_weights(c[:, 0], r, None, ck)
    ld = criterion(out, targets, w)
    total = sum(ld[k] * criterion.weight_dict[k] for k in ld if k in criterion.weight_dict)
    opt.zero_grad()
    total.backward()
    buckets.all_reduce()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
agd.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agd.barrier()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ker = [(e.name, e.time_range.start, e.time_range.end) for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    nccl = [k for k in ker if "nccl" in k[0].lower()]
    comp = [k for k in ker if "nccl" not in k[0].lower()]
    t0 = min(k[1] for k in ker)
    span = max(k[2] for k in ker) - t0
    print(f"# one training step, 2 x 150k voxels per GPU, {world} GPUs, rank 0: {len(comp)} compute kernels, {len(nccl)} NCCL kernels, "
          f"step span {span / 1e3:.2f} ms (under the profiler)")
    tot_n = tot_o = 0.0
    for name, s, e in nccl:
        over, names = 0.0, {}
        for cn, cs, ce in comp:
            o = min(e, ce) - max(s, cs)
            if o > 0:
                over += o
                key = cn.split("(")[0][-48:]
                names[key] = names.get(key, 0.0) + o
        tot_n += e - s
        tot_o += min(over, e - s)
        top = ", ".join(f"{k} {v / 1e3:.2f} ms" for k, v in sorted(names.items(), key=lambda kv: -kv[1])[:4])
        print(f"{name.split('(')[0][:40]:40s} start {(s - t0) / 1e3:8.2f} ms  dur {(e - s) / 1e3:6.2f} ms  concurrent compute {min(over, e - s) / 1e3:6.2f} ms  [{top}]")
    print(f"# NCCL kernel time {tot_n / 1e3:.2f} ms, of which {tot_o / 1e3:.2f} ms ({100 * tot_o / max(tot_n, 1e-9):.0f} %) ran concurrently with this library's kernels")
if world > 1:
    torch.distributed.destroy_process_group()
