#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s40
for n in 2 1 3; do timeout 900 python bench.py --steps 12 --no-parity --pipeline $n > gpurun_out/${S}_bench_p$n.json 2> gpurun_out/${S}_p$n.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${S}_bench_p$n.json"))
    print("pipeline $n", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("pipeline $n failed", e)
PY
tail -n 3 gpurun_out/${S}_p$n.err
done
