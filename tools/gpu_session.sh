#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s39
for n in 3 4 8; do AG3D_DEC_STREAMS=$n timeout 900 python bench.py --steps 12 --no-parity > gpurun_out/${S}_bench_s$n.json 2> gpurun_out/${S}_s$n.err; python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_s$n.json"))
print("streams $n", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
PY
done
AG3D_DEC_STREAMS=4 timeout 900 python bench.py --steps 12 --no-parity --batch 1 > gpurun_out/${S}_bench_b1.json 2> gpurun_out/${S}_b1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_b1.json"))
print("b1", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
PY
