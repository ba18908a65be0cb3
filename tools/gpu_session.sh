#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s31
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --no-parity > gpurun_out/${S}_bench_n2.json 2> gpurun_out/${S}_n2.err
timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 5 --workload train --batch 4 > gpurun_out/${S}_train_n2_overlap.json 2> gpurun_out/${S}_train_n2a.err
AG3D_NO_OVERLAP=1 timeout 900 $TR --master-port 29543 bench.py --gpus 2 --steps 5 --workload train --batch 4 > gpurun_out/${S}_train_n2_after.json 2> gpurun_out/${S}_train_n2b.err
timeout 900 $TR --master-port 29544 bench.py --gpus 2 --steps 4 --workload train --batch 4 --voxels 500000 > gpurun_out/${S}_train_c4_n2.json 2> gpurun_out/${S}_train_c4_n2.err
for f in bench_n2 train_n2_overlap train_n2_after train_c4_n2; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${S}_$f.json"))
    print("$f", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["config"].get("gradient_exchange"))
except Exception as e:
    print("$f no json:", e)
PY
done
tail -n 2 gpurun_out/${S}_*n2*.err | tail -n 16
