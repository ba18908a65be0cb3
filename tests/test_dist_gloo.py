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
