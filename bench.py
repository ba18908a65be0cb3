#!/usr/bin/env python
"""bench.py — scenes/s forward of the AGILE3D hot path (backbone + click-query decoder) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric, SURVEY.md §8(d) "headline"): synthetic ScanNet-shape scenes of ~150k voxels
(2 cm), 5 objects x 2 clicks = 10 clicks -> 20 click queries, eval-mode ``forward_backbone`` + ``forward_mask``
(kernel maps are rebuilt for every scene, as the reference does per scene).  One step = one batch of B scenes.
One JSON line on stdout (rank 0).  Weak scaling: every rank runs its own B scenes per step, no collectives on
the data path (inference shards by scene).  `--pipeline P` (default 2): P batches in flight on
alternating streams, as a serving loop would run them; every step is issued and completed inside the timed region;
`--pipeline 1` runs the steps back to back.  The `e2e` value repeats the measurement with pinned HOST buffers: the
inputs of every step are copied to the device and its logits back, on side streams, inside the timed region.

`--impl reference` times the CPU restatement of the reference path (oracle/, fp32, all host threads): the
reference's real CPU path needs MinkowskiEngine, which cannot be installed here.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

TARGET_VOXELS = 150000
VOXEL = 0.02
N_OBJ, CLICKS_PER_OBJ, N_BG = 5, 2, 0
METRIC = "scenes/sec forward (150k voxels, 10 click queries)"
WORKLOAD = "150k-voxel synthetic ScanNet-shape scene @2cm, 5 objects x 2 clicks (20 queries), eval forward_backbone+forward_mask"


def make_inputs(n_scenes, seed0, target=TARGET_VOXELS):
    from agile3d_b200.scenes import make_clicks, make_scene
    scenes = []
    for i in range(n_scenes):
        sc = make_scene(target, VOXEL, seed=seed0 + i)
        clicks, times, _ = make_clicks(sc, N_OBJ, CLICKS_PER_OBJ, N_BG, seed=seed0 + i)
        scenes.append((sc, clicks, times))
    return scenes


def collate(batch):
    """list of (scene, clicks, times) -> pinned host tensors + click lists (the reference's collation_fn)."""
    from agile3d_b200 import utils
    coords = utils.batched_coordinates([s["coords"] for s, _, _ in batch])
    feats = torch.from_numpy(np.concatenate([s["feats"] for s, _, _ in batch], 0))
    raw = torch.from_numpy(np.concatenate([s["raw_coords"] for s, _, _ in batch], 0))
    return coords, feats, raw, [c for _, c, _ in batch], [t for _, _, t in batch]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f).get("bf16_tflops_sustained", 1400.0))
    return 1400.0


def ncu_summary():
    """per-launch DRAM bytes / tensor-pipe activity of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_dominant.json, written by tools/ncu_summary.py from the .ncu-rep of the same command)"""
    p = os.path.join(ROOT, "profiles", "ncu_dominant.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_run(scene_tuple, model=None, repeats=1):
    """fp32 CPU oracle (restated reference path) on one scene; returns seconds per scene (median)."""
    from agile3d_b200 import utils
    from agile3d_b200.weights import default_args, synth_state_dict
    from oracle import me_ref
    from oracle.agile3d_ref import build_ref_model
    torch.set_num_threads(os.cpu_count() or 1)
    if model is None:
        model = build_ref_model(default_args()).eval()
        model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
    sc, clicks, times = scene_tuple
    coords = utils.batched_coordinates([sc["coords"]])
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            x = me_ref.SparseTensor(coordinates=coords, features=torch.from_numpy(sc["feats"]))
            h = model.forward_backbone(x, torch.from_numpy(sc["raw_coords"]))
            model.forward_mask(*h, [clicks], [times])
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), model


def crop_scene(scene_tuple, n_keep):
    """Spatially contiguous crop (lowest x first) with clicks re-chosen inside the crop."""
    from agile3d_b200.scenes import make_clicks
    sc, _, _ = scene_tuple
    order = np.argsort(sc["raw_coords"][:, 0], kind="stable")[:n_keep]
    order.sort()
    sub = dict(sc)
    for k in ("coords", "raw_coords", "feats", "labels"):
        sub[k] = np.ascontiguousarray(sc[k][order])
    try:
        clicks, times, _ = make_clicks(sub, N_OBJ, CLICKS_PER_OBJ, 0, seed=1)
    except AssertionError:                              # tiny crops may hold fewer boxes: click the shell instead
        rows = np.linspace(0, n_keep - 1, N_OBJ * CLICKS_PER_OBJ).astype(int).tolist()
        clicks = {"0": []}
        times = {"0": []}
        for j in range(N_OBJ):
            clicks[str(j + 1)] = rows[2 * j:2 * j + 2]
            times[str(j + 1)] = [2 * j, 2 * j + 1]
    return sub, clicks, times


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    full = make_inputs(1, 2000)[0]
    n_full = full[0]["coords"].shape[0]
    # calibrate on a 15k-voxel crop, then size the per-step sample so the whole run stays within ~150 s
    t_cal, model = cpu_reference_run(crop_scene(full, 15000))
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_keep = int(min(n_full, max(15000, 15000 * budget / max(t_cal, 1e-3))))
    sample = full if n_keep >= n_full else crop_scene(full, n_keep)
    n_s = sample[0]["coords"].shape[0]
    for _ in range(args.warmup):
        cpu_reference_run(sample, model)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(sample, model)
    dt = time.perf_counter() - t0
    # one step covers n_s / n_full of a headline scene; cost is ~linear in voxels
    value = args.steps * (n_s / n_full) / dt
    sample_desc = (f"{n_s}-voxel crop of a {n_full}-voxel scene per step (scaled by voxel count); "
                   "oracle/ fp32 torch restatement of MinkowskiEngine + reference decoder, not ME's own kernels")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": "scenes/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import agile3d_b200
    from agile3d_b200 import ops
    from agile3d_b200.weights import default_args, synth_state_dict
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    from agile3d_b200 import dist as agd
    dist_on = world > 1
    agd.init_from_env(backend="nccl", device=dev)       # no-op for a single process
    model = agile3d_b200.build_model(default_args()).eval()
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
    model = model.to(dev)
    B = args.batch
    n_pool = 2                                                     # distinct batches, rotated every step
    scenes = make_inputs(B * n_pool, 2000 + 100 * rank)
    host = [collate(scenes[i * B:(i + 1) * B]) for i in range(n_pool)]
    host = [(c.pin_memory(), f.pin_memory(), r.pin_memory(), ck, tm) for c, f, r, ck, tm in host]
    resident = [(c.to(dev), f.to(dev), r.to(dev), ck, tm) for c, f, r, ck, tm in host]
    n_vox = [int(c.shape[0]) for c, *_ in host]

    # Batches are independent, so a serving loop keeps two of them in flight: consecutive steps run on alternating streams
    # and the latency-bound phases of one batch (coarse U-Net levels, click-query kernels, coordinate maps) execute under the
    # bandwidth- / tensor-bound phases of the other.  Every step is still issued and completed inside the timed region
    # (fork_pipe / join_pipe bracket it on the timing stream).  --pipeline 1 runs the steps back to back on one stream.
    pipe = [torch.cuda.Stream(device=dev) for _ in range(max(args.pipeline, 1))]
    piped = [args.pipeline > 1]

    def step_stream(i):
        return torch.cuda.stream(pipe[i % len(pipe)]) if piped[0] else contextlib.nullcontext()

    def fork_pipe():
        if piped[0]:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for st in pipe:
                st.wait_event(ev)

    def join_pipe():
        if piped[0]:
            for st in pipe:
                torch.cuda.current_stream().wait_stream(st)

    def step_resident(i):
        c, f, r, ck, tm = resident[i % n_pool]
        with step_stream(i):
            x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
            h = model.forward_backbone(x, raw_coordinates=r)
            return model.forward_mask(*h, click_idx=ck, click_time_idx=tm)

    out_host = {}
    # End to end through the public API with HOST buffers: every step copies its inputs from pinned host memory and its
    # logits back.  The copies run on their own streams, double-buffered, as a serving loop would: the inputs of step
    # i+1 travel while step i computes, the logits of step i while step i+1 computes; every copy is issued inside the
    # timed region and the region ends only when the last logits have landed (finish_e2e).
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    staged = {}

    def stage_inputs(i):
        c, f, r, ck, tm = host[i % n_pool]
        with torch.cuda.stream(h2d_stream):
            t = (c.to(dev, non_blocking=True), f.to(dev, non_blocking=True), r.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(h2d_stream)
        staged[i] = (t, ev)

    def step_e2e(i):
        if i not in staged:
            stage_inputs(i)
        (cd, fd, rd), ev = staged.pop(i)
        with step_stream(i):
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in (cd, fd, rd):
                t.record_stream(cur)
            stage_inputs(i + 1)                                    # next step's inputs travel under this step's kernels
            _, _, _, ck, tm = host[i % n_pool]
            x = agile3d_b200.SparseTensor(coordinates=cd, features=fd, device=dev)
            h = model.forward_backbone(x, raw_coordinates=rd)
            out = model.forward_mask(*h, click_idx=ck, click_time_idx=tm)
            done = torch.cuda.Event()
            done.record(cur)
        d2h_stream.wait_event(done)
        with torch.cuda.stream(d2h_stream):
            for b, p in enumerate(out["pred_masks"]):              # the caller reads the logits (eval_multi_obj.py:124-125)
                key = (i % n_pool, b)
                if key not in out_host:
                    out_host[key] = torch.empty(p.shape, dtype=p.dtype, pin_memory=True)
                out_host[key].copy_(p, non_blocking=True)
                p.record_stream(d2h_stream)
        return out

    def finish_e2e():
        join_pipe()
        torch.cuda.current_stream().wait_stream(d2h_stream)
        torch.cuda.current_stream().wait_stream(h2d_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            agd.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork_pipe()                                                # the step streams start after e0 ...
        for i in range(steps):
            fn(i)
        (finish or join_pipe)()                                    # ... and e1 follows the last kernel of every stream
        e1.record()
        barrier()
        return agd.max_over_ranks(e0.elapsed_time(e1), device=dev)     # slowest rank = the job's time

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    join_pipe()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.kernel_launches()
    ms = timed(step_resident, args.steps)
    launches = ops.kernel_launches() - l0
    for i in range(2):
        step_e2e(i)
    finish_e2e()
    torch.cuda.synchronize()
    staged.clear()                                                 # step 0 of the timed region copies its own inputs
    ms_e2e = timed(step_e2e, args.steps, finish_e2e)
    staged.clear()
    clocks = sampler.stop() if sampler else None

    # per-family CUDA-event pass for the roofline (same workload, separate from the headline timing)
    fam = None
    if rank == 0:
        torch.cuda.synchronize()
        piped[0] = False                                           # one batch, one stream: clean per-kernel event times
        saved_streams, model.decoder_streams = model.decoder_streams, 0
        prof = ops.Profiler()
        ops.set_profiler(prof)
        for i in range(2):
            step_resident(i)
        ops.set_profiler(None)
        fam = prof.summary()
        piped[0] = args.pipeline > 1
        model.decoder_streams = saved_streams

    parity = None
    if rank == 0 and not args.no_parity:
        # measured parity of the timed workload itself: first scene of the first timed batch against the fp64 CPU oracle
        # (oracle/compare.py: layers after a discrete label decision are compared on the same decisions)
        from oracle.agile3d_ref import build_ref_model
        from oracle.compare import decision_forced_errors, rel_err
        sc0, ck0, tm0 = scenes[0]
        c0 = np.concatenate([np.zeros((sc0["coords"].shape[0], 1), np.int32), sc0["coords"]], 1)
        x0 = agile3d_b200.SparseTensor(coordinates=torch.from_numpy(c0), features=torch.from_numpy(sc0["feats"]), device=dev)
        h0 = model.forward_backbone(x0, raw_coordinates=torch.from_numpy(sc0["raw_coords"]).to(dev))
        o0 = model.forward_mask(*h0, click_idx=[ck0], click_time_idx=[tm0])
        layers = [a["pred_masks"] for a in o0["aux_outputs"]] + [o0["pred_masks"]]
        ref = build_ref_model(default_args()).eval()
        ref.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed=5))
        r = decision_forced_errors(ref.double(), c0, sc0["feats"], sc0["raw_coords"], [ck0], [tm0], layers)
        parity = {"checker": "oracle/ fp64 CPU restatement, scene 0 of the timed batch",
                  "backbone_features_rel_err": rel_err(h0[0].F.cpu().numpy(), r["pcd"].F.numpy()),
                  "mask_logits_rel_err_per_layer": r["forced"], "vs_free_running_oracle": r["free"],
                  "differing_label_decisions": r["flips"], "not_near_ties": r["bad_flips"], "tolerance": 1e-3}

    agd.barrier()
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    total_scenes = world * B * args.steps
    value = total_scenes / (ms / 1e3)
    e2e_value = total_scenes / (ms_e2e / 1e3)
    fam_rows = {}
    for name, f in fam.items():
        gbs = f["bytes"] / (f["ms"] / 1e3) / 1e9 if f["ms"] > 0 else 0.0
        fam_rows[name] = {"launches_per_step": f["launches"] // 2, "ms_per_step": round(f["ms"] / 2, 4),
                          "algorithmic_GB_per_step": round(f["bytes"] / 2 / 1e9, 4), "GBps": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / peak, 4),
                          "TFLOPs": round(f["flops"] / (f["ms"] / 1e3) / 1e12, 3) if f["ms"] > 0 else 0.0}
    tf_peak = measured_tflops()
    for name, f in fam.items():
        # the other roof: tensor pipe.  The sparse conv computes every product three times (bf16x3 = fp32 parity on bf16
        # tensor cores), the decoder GEMMs likewise; "binding" names the roof that is closer.
        if f["flops"]:
            t_tensor = 3.0 * f["flops"] / (tf_peak * 1e12) * 1e3
            t_hbm = f["bytes"] / (peak * 1e9) * 1e3
            fam_rows[name].update({"hbm_bound_ms_per_step": round(t_hbm / 2, 4), "tensor_bound_ms_per_step": round(t_tensor / 2, 4),
                                   "binding": "tensor" if t_tensor > t_hbm else "hbm",
                                   "frac_of_binding_bound": round(max(t_tensor, t_hbm) / f["ms"], 4)})
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
    ncu = ncu_summary()
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu.get("dram_bytes_per_launch"),
                "tensor_pipe_pct": ncu.get("tensor_pipe_pct"), "ncu_source": ncu.get("source"), "peak_source": peak_src,
                "per_launch_algorithmic_bytes": int(d["bytes"] / max(d["launches"], 1)),
                "per_launch_ms": round(d["ms"] / max(d["launches"], 1), 5),
                "note": "the dominant family is the sparse convolution; at fp32 parity (bf16x3) its binding roof is the tensor "
                        "pipe, see families.spconv.binding / frac_of_binding_bound",
                "tensor_peak_TFLOPs": tf_peak, "families": fam_rows}
    # CPU baseline on a bounded sample: one crop sized for ~10-20 s of host work
    full = scenes[0]
    t_cal, ref_model = cpu_reference_run(crop_scene(full, 10000))
    n_keep = int(min(full[0]["coords"].shape[0], max(10000, 10000 * 12.0 / max(t_cal, 1e-3))))
    sample = full if n_keep >= full[0]["coords"].shape[0] else crop_scene(full, n_keep)
    t_cpu, _ = cpu_reference_run(sample, ref_model)
    n_s, n_full = sample[0]["coords"].shape[0], full[0]["coords"].shape[0]
    cpu_value = (n_s / n_full) / t_cpu
    h2d = int(np.mean([n * (16 + 12 + 12) for n in n_vox]))
    d2h = int(np.mean([n * (1 + N_OBJ) * 4 for n in n_vox]))
    line = {
        "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scenes_per_step_per_gpu": B, "voxels_per_step_per_gpu": int(np.mean(n_vox)),
                   "l2_policy": f"rotating pool of {n_pool} distinct batches; per-step activations (~1.7 GB/scene) exceed the 126 MB L2",
                   "batches_in_flight": max(args.pipeline, 1),
                   "decoder_streams": int(model.decoder_streams),
                   "spconv_algo": {0: "auto", 1: "simt_fp32", 2: "tcgen05_bf16x3"}[model.backbone.algo]},
        "e2e": {"value": e2e_value, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "copies": "pinned host buffers, every step; H2D of step i+1 and D2H of step i on side streams under the kernels "
                          "of the neighbouring step, all inside the timed region"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "parity": parity,
        "cpu_baseline": {"value": cpu_value, "unit": "scenes/s", "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": f"{n_s}-voxel crop of a {n_full}-voxel scene, 1 run, scaled by voxel count; "
                                   "oracle/ fp32 torch restatement (MinkowskiEngine itself is not installable here)"},
    }
    emit(line)
    if dist_on:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------ training step
TRAIN_METRIC = "scenes/sec train step (150k voxels, 10 click queries; forward+backward+criterion+clip+AdamW)"
TRAIN_WORKLOAD = ("BASELINE configs[2]: batch of 150k-voxel synthetic scenes @2cm, 5 objects x 2 clicks (20 queries), "
                  "train-mode forward_backbone+forward_mask, SetCriterion with click loss weights, backward, "
                  "gradient all-reduce (N>1), clip_grad_norm 0.1 + AdamW")


def run_train(args, rank, world, local_rank):
    """`--workload train`: one step = the reference's engine.py:119-150 on a batch of B scenes per GPU (weak scaling:
    every rank has its own B scenes; the only collective is the bucketed gradient all-reduce)."""
    import agile3d_b200
    from agile3d_b200 import dist as agd
    from agile3d_b200 import ops
    from agile3d_b200.optim import FlatAdamW, GradBuckets
    from agile3d_b200.scenes import make_clicks, make_scene
    from agile3d_b200.weights import default_args, synth_state_dict
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist_on = world > 1
    agd.init_from_env(backend="nccl", device=dev)
    margs = default_args()
    model = agile3d_b200.build_model(margs)
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
    model = model.to(dev).train()
    criterion = agile3d_b200.build_criterion(margs)
    opt = FlatAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, max_norm=0.1)        # main.py:62-69,125
    buckets = GradBuckets(opt, n_buckets=6)
    # the exchange of a stage's gradients starts inside the backbone backward as soon as the stage is done
    # (AG3D_NO_OVERLAP=1: exchange after the backward, for the comparison in profiles/)
    if os.environ.get("AG3D_NO_OVERLAP", "0") != "1":
        buckets.attach(model)
    target_voxels = args.voxels or TARGET_VOXELS
    B, n_pool = args.batch, 2
    host = []
    for pidx in range(n_pool):
        batch, targets = [], []
        for i in range(B):
            seed = 2000 + 100 * rank + pidx * B + i
            sc = make_scene(target_voxels, VOXEL, seed=seed)
            clicks, times, lab = make_clicks(sc, N_OBJ, CLICKS_PER_OBJ if target_voxels < 400000 else 3, 0, seed=seed)
            batch.append((sc, clicks, times))
            targets.append(torch.from_numpy(lab.astype(np.int32)))
        c, f, r, ck, tm = collate(batch)
        host.append((c.pin_memory(), f.pin_memory(), r.pin_memory(), ck, tm, [t.pin_memory() for t in targets]))
    resident = [(c.to(dev), f.to(dev), r.to(dev), ck, tm, [t.to(dev) for t in tg]) for c, f, r, ck, tm, tg in host]
    n_vox = [int(c.shape[0]) for c, *_ in host]
    loss_host = torch.empty(1, dtype=torch.float32, pin_memory=True)

    def one(c, f, r, ck, tm, tg):
        x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
        h = model.forward_backbone(x, raw_coordinates=r)
        out = model.forward_mask(*h, click_idx=ck, click_time_idx=tm)
        w = agile3d_b200.cal_click_loss_weights(c[:, 0], r, None, ck)
        ld = criterion(out, tg, w)
        total = sum(ld[k] * criterion.weight_dict[k] for k in ld if k in criterion.weight_dict)
        opt.zero_grad()
        total.backward()
        buckets.all_reduce()
        opt.step()
        return total

    def step_resident(i):
        return one(*resident[i % n_pool])

    def step_e2e(i):
        c, f, r, ck, tm, tg = host[i % n_pool]
        total = one(c.to(dev, non_blocking=True), f.to(dev, non_blocking=True), r.to(dev, non_blocking=True), ck, tm,
                    [t.to(dev, non_blocking=True) for t in tg])
        loss_host.copy_(total.detach().reshape(1), non_blocking=True)      # the trainer reads the loss (engine.py:138)

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            agd.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return agd.max_over_ranks(e0.elapsed_time(e1), device=dev)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.kernel_launches()
    ms = timed(step_resident, args.steps)
    launches = ops.kernel_launches() - l0
    for i in range(n_pool):                      # one untimed pass per pool entry (allocator warm-up of the staging copies)
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    # per-family pass: every rank runs the step (it contains the gradient all-reduce), rank 0 records it
    fam = None
    prof = ops.Profiler() if rank == 0 else None
    ops.set_profiler(prof)
    step_resident(0)
    ops.set_profiler(None)
    if rank == 0:
        fam = prof.summary()
    agd.barrier()
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    total_scenes = world * B * args.steps
    fam_rows = {}
    for name, f in fam.items():
        gbs = f["bytes"] / (f["ms"] / 1e3) / 1e9 if f["ms"] > 0 else 0.0
        fam_rows[name] = {"launches_per_step": f["launches"], "ms_per_step": round(f["ms"], 4),
                          "algorithmic_GB_per_step": round(f["bytes"] / 1e9, 4), "GBps": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / peak, 4),
                          "TFLOPs": round(f["flops"] / (f["ms"] / 1e3) / 1e12, 3) if f["ms"] > 0 else 0.0}
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
    line = {
        "metric": TRAIN_METRIC, "value": total_scenes / (ms / 1e3), "unit": "scenes/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD if target_voxels < 400000 else TRAIN_WORKLOAD.replace(
                       "configs[2]: batch of 150k-voxel", "configs[3]: batch of 500k-voxel (S3DIS-shape)").replace(
                       "5 objects x 2 clicks (20 queries)", "5 objects x 3 clicks (25 queries)"),
                   "scenes_per_step_per_gpu": B, "gradient_exchange": "overlapped with the backward (GradBuckets.attach)"
                   if buckets.model is not None else "after the backward",
                   "voxels_per_step_per_gpu": int(np.mean(n_vox)),
                   "l2_policy": f"rotating pool of {n_pool} distinct batches; activations exceed the 126 MB L2"},
        "e2e": {"value": total_scenes / (ms_e2e / 1e3), "unit": "scenes/s",
                "h2d_bytes_per_step": int(np.mean([n * (16 + 12 + 12 + 4) for n in n_vox])), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "families": fam_rows},
        "cpu_baseline": None,
    }
    emit(line)
    if dist_on:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ click loop (configs[4])
CLICK_METRIC = "scenes/sec iterative-click evaluation (80k voxels @5cm, 10 objects, 200 clicks; eval_multi_obj.py protocol)"
CLICK_WORKLOAD = ("BASELINE configs[4]: KITTI-360-shape outdoor scan ~80k voxels @5cm, 10 objects; forward_backbone once, then the "
                  "click loop of eval_multi_obj.py:118-167 (forward_mask, prediction update, full-resolution IoU, simulated next "
                  "click) until 20 clicks per object: ~192 rounds, click queries 20 -> 210")


def run_clickloop(args, rank, world, local_rank):
    """One step = the whole interactive protocol on one scene per GPU (scenes shard across ranks, no collective)."""
    import copy
    import random

    import agile3d_b200
    from agile3d_b200 import dist as agd
    from agile3d_b200 import interactive, ops
    from agile3d_b200.scenes import make_clicks, make_scene
    from agile3d_b200.weights import default_args, synth_state_dict
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist_on = world > 1
    agd.init_from_env(backend="nccl", device=dev)
    model = agile3d_b200.build_model(default_args()).eval()
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=5))
    model = model.to(dev)
    K, max_per_obj = 10, 20                                              # eval_multi_obj.py --max_num_clicks 20
    scenes = []
    for i in range(2):
        sc = make_scene(80000, 0.05, seed=5000 + 100 * rank + i, outdoor=True)
        _, _, lab = make_clicks(sc, K, 1, 0, seed=5000 + i)
        scenes.append((sc, torch.from_numpy(np.minimum(lab, K).astype(np.int64))))
    host = []
    for sc, lab in scenes:
        coords = agile3d_b200.utils.batched_coordinates([sc["coords"]])
        lab_full = lab[sc["inverse_map"]]
        host.append((coords.pin_memory(), torch.from_numpy(sc["feats"]).pin_memory(), torch.from_numpy(sc["raw_coords"]).pin_memory(),
                     lab.pin_memory(), lab_full.pin_memory(), sc["inverse_map"].pin_memory()))
    rounds_seen, nq_seen = [], []

    def protocol(i, resident):
        c, f, r, lab, lab_full, inv = resident[i % len(resident)]
        if not c.is_cuda:
            c, f, r, lab, lab_full, inv = (t.to(dev, non_blocking=True) for t in (c, f, r, lab, lab_full, inv))
        random.seed(i)
        x = agile3d_b200.SparseTensor(coordinates=c, features=f, device=dev)
        h = model.forward_backbone(x, raw_coordinates=r)
        click_idx = {str(o): [] for o in range(K + 1)}
        click_time = copy.deepcopy(click_idx)
        n_clicks, rounds, iou = 0, 0, None
        nv = c.shape[0]
        while n_clicks <= K * max_per_obj:
            if n_clicks == 0:
                pred = ops.click_pred(None, nv, K + 1, torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
            else:
                out = model.forward_mask(*h, click_idx=[click_idx], click_time_idx=[click_time])
                rows = torch.tensor([v for o in range(K + 1) for v in click_idx[str(o)]], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
                objs = torch.tensor([o for o in range(K + 1) for _ in click_idx[str(o)]], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
                pred = ops.click_pred(out["pred_masks"][0], nv, K + 1, rows, objs)
                nq_seen.append(10 + n_clicks)
            (iou, _), (new, _, _, new_t) = interactive.iou_and_simulated_clicks(pred, lab_full, inv, lab, r, n_clicks, n_obj=K + 1)
            if new is not None:
                click_idx, click_time = interactive.extend_clicks(click_idx, click_time, new, new_t)
            n_clicks += K if n_clicks == 0 else 1
            rounds += 1
        rounds_seen.append(rounds)
        return iou

    resident = [tuple(t.to(dev) for t in hs) for hs in host]

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            agd.barrier()
            torch.cuda.synchronize()

    def timed(res, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            protocol(i, res)
        e1.record()
        barrier()
        return agd.max_over_ranks(e0.elapsed_time(e1), device=dev)

    steps = max(1, min(args.steps, 5))
    for i in range(max(1, min(args.warmup, 2))):
        protocol(i, resident)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.kernel_launches()
    ms = timed(resident, steps)
    launches = ops.kernel_launches() - l0
    ms_e2e = timed(host, steps)
    clocks = sampler.stop() if sampler else None
    agd.barrier()
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    rounds = rounds_seen[-1]
    n_v = int(host[0][0].shape[0])
    line = {
        "metric": CLICK_METRIC, "value": world * steps / (ms / 1e3), "unit": "scenes/s", "n_gpus": world, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CLICK_WORKLOAD, "voxels": n_v, "rounds_per_scene": rounds, "ms_per_round": ms / steps / rounds,
                   "click_queries": [min(nq_seen), max(nq_seen)], "host_syncs_per_round": 1,
                   "l2_policy": "two scenes rotated; a round streams the 41 MB voxel features + 41 MB encodings 6 times"},
        "e2e": {"value": world * steps / (ms_e2e / 1e3), "unit": "scenes/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": n_v * (16 + 12 + 12 + 8) + int(host[0][4].shape[0]) * 16, "d2h_bytes_per_step": rounds * (64 * 4 + 11 * 24)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "cpu_baseline": None,
    }
    emit(line)
    if dist_on:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """Exactly one JSON line on the real stdout (libraries such as NCCL may print banners to fd 1)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)           # keep the real stdout for the JSON line ...
    os.dup2(2, 1)                      # ... and send everything else written to fd 1 to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="scenes per step per GPU (BASELINE configs[2] batches 8 scenes)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="forward", choices=["forward", "train", "c2", "clickloop"],
                    help="forward = the headline metric (default); train = BASELINE configs[2]/[3] training step; c2 = configs[1] "
                         "(150k voxels, 5 clicks, single object, batch 1); clickloop = configs[4] (80k voxels @5cm, the iterative-click "
                         "protocol of eval_multi_obj.py end to end)")
    ap.add_argument("--pipeline", type=int, default=2, help="forward workloads: batches in flight on alternating streams (1 = back to back)")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp64 oracle comparison of the timed workload")
    ap.add_argument("--voxels", type=int, default=0, help="train workload: voxels per scene (default 150k; 500000 = configs[3])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if args.workload == "train":
        run_train(args, rank, world, local_rank)
        return
    if args.workload == "clickloop":
        run_clickloop(args, rank, world, local_rank)
        return
    if args.workload == "c2":
        global N_OBJ, CLICKS_PER_OBJ, N_BG, WORKLOAD, METRIC
        N_OBJ, CLICKS_PER_OBJ, N_BG = 1, 3, 2
        METRIC = "scenes/sec forward (150k voxels, 5 clicks, single object)"
        WORKLOAD = "BASELINE configs[1]: 150k-voxel synthetic ScanNet-shape scene @2cm, 1 object, 3 fg + 2 bg clicks (15 queries), eval forward_backbone+forward_mask"
        if args.batch == 8:
            args.batch = 1
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
