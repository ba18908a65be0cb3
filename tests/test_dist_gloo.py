"""N > 1 path on CPU: two gloo ranks shard the scenes, run the (oracle) forward on their shard, and the job
throughput bookkeeping (SUM of units / MAX of times) agrees on both ranks.  No collective touches the data path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agile3d_b200 import dist as agd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = agd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    mine = agd.shard_scenes(5, r, w)
    # every rank "processes" its scenes (a deterministic stand-in cost: rank 1 is slower)
    seconds = 0.5 + rank
    agd.barrier()
    thr = agd.job_throughput(len(mine), seconds)
    tmax = agd.max_over_ranks(seconds)
    total = agd.sum_over_ranks(len(mine))
    out[rank] = (mine, thr, tmax, total)
    dist.destroy_process_group()


def test_two_rank_sharding_and_throughput():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert sorted(res[0][0] + res[1][0]) == [0, 1, 2, 3, 4]          # every scene exactly once
    assert abs(len(res[0][0]) - len(res[1][0])) <= 1
    for r in (0, 1):
        assert res[r][2] == 1.5 and res[r][3] == 5.0
        assert abs(res[r][1] - 5.0 / 1.5) < 1e-12


def test_shard_and_balance_properties():
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 4, 8):
            parts = [agd.shard_scenes(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    sizes = [500_000, 80_000, 150_000, 150_000, 20_000, 300_000, 90_000, 150_000]
    owned = agd.balance_by_voxels(sizes, 4)
    assert sorted(i for o in owned for i in o) == list(range(8))
    loads = [sum(sizes[i] for i in o) for o in owned]
    assert max(loads) <= 500_000                                       # the largest scene bounds the makespan here
    assert agd.max_over_ranks(3.0) == 3.0 and agd.sum_over_ranks(2.0) == 2.0   # single process: identity


def _dp_worker(rank, world, port, out):
    """Data-parallel training exchange: every rank holds different gradients in FlatAdamW's flat
    buffer; GradBuckets sums + averages them bucket by bucket; the clip + AdamW step (C-ABI ops replaced by their
    contract emulation on CPU) then leaves identical parameters on every rank."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import emulate
    import agile3d_b200.ops as ops
    from agile3d_b200.optim import FlatAdamW, GradBuckets
    for name in ("grad_norm", "adamw_step"):
        setattr(ops, name, getattr(emulate, name))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    agd.init_from_env(backend="gloo")
    torch.manual_seed(0)                                     # same initial parameters everywhere
    params = [torch.nn.Parameter(torch.randn(37, 5)), torch.nn.Parameter(torch.randn(101)), torch.nn.Parameter(torch.randn(3, 3, 3))]
    opt = FlatAdamW(params, lr=1e-2, weight_decay=1e-2, max_norm=0.1)
    buckets = GradBuckets(opt, n_buckets=4)
    opt.zero_grad()
    torch.manual_seed(100 + rank)                            # different data -> different gradients
    loss = sum((p * torch.randn_like(p)).sum() for p in params)
    loss.backward()
    local = opt.flat_g.clone()
    buckets.all_reduce()
    norm = opt.step()
    out[rank] = (local.numpy(), opt.flat_g.clone().numpy(), opt.flat_p.clone().numpy(), float(norm))
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_and_step():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    mean = (res[0][0] + res[1][0]) / 2
    for r in (0, 1):
        assert abs(res[r][1] - mean).max() < 1e-6            # averaged gradients everywhere
    assert abs(res[0][2] - res[1][2]).max() == 0.0           # identical parameters after the step
    assert res[0][3] == res[1][3]
