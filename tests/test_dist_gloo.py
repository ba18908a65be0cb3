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


def _overlap_worker(rank, world, port, out):
    """The OVERLAPPED exchange (GradBuckets.attach): the backbone backward hands its gradients to the buckets stage by
    stage; after finish() the flat buffer must hold the mean of the ranks' gradients for EVERY parameter (the stage
    slices tile the buffer, nothing is exchanged twice or not at all)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import emulate
    import agile3d_b200
    import agile3d_b200.ops as ops
    from agile3d_b200.optim import FlatAdamW, GradBuckets
    from agile3d_b200.weights import default_args, synth_state_dict
    from helpers import load_golden
    for name in emulate.ALL:
        setattr(ops, name, getattr(emulate, name))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))                 # two ranks share the host cores
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    agd.init_from_env(backend="gloo")
    g = load_golden("train_g1200_k2")
    m = agile3d_b200.build_model(default_args())
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=g["wseed"]))
    m.train()
    criterion = agile3d_b200.build_criterion(default_args())
    opt = FlatAdamW(m.parameters(), lr=1e-3, weight_decay=0.0, max_norm=0.1)
    buckets = GradBuckets(opt, n_buckets=4)
    coords = torch.from_numpy(g["coords"])
    feats = torch.from_numpy(g["feats"]) * (1.0 + 0.5 * rank)                 # different data on every rank
    raw = torch.from_numpy(g["raw_coords"])
    targets = [torch.from_numpy(g["targets"])]
    weights = agile3d_b200.cal_click_loss_weights(coords[:, 0], raw, torch.cat(targets), g["clicks"])

    def backward():
        opt.zero_grad()
        x = agile3d_b200.SparseTensor(coordinates=coords, features=feats)
        out_ = m.forward_mask(*m.forward_backbone(x, raw), g["clicks"], g["times"])
        ld = criterion(out_, targets, weights)
        sum(ld[k] * criterion.weight_dict[k] for k in ld if k in criterion.weight_dict).backward()

    backward()
    local = opt.flat_g.clone()                                                # plain autograd accumulation, no exchange
    for mod in m.modules():                                                   # same BatchNorm state for the second pass
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.reset_running_stats()
    buckets.attach(m)
    backward()
    covered = sorted(buckets.launched)
    buckets.all_reduce()
    out[rank] = (local.numpy(), opt.flat_g.clone().numpy(), covered, opt.flat_g.numel())
    dist.destroy_process_group()


def test_two_rank_overlapped_exchange_covers_every_parameter():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_overlap_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    mean = (res[0][0] + res[1][0]) / 2
    scale = abs(mean).max()
    for r in (0, 1):
        assert abs(res[r][1] - mean).max() < 1e-5 * scale
        cov, n = res[r][2], res[r][3]
        assert cov[0][0] == 0 and cov[-1][1] == n and all(a[1] == b[0] for a, b in zip(cov[:-1], cov[1:])), cov
        assert len(cov) >= 9                                                  # tail + 4 decoder + 4 encoder stages (+ stem)
