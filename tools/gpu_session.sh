#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s45
timeout 900 python bench.py > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --steps 10 --batch 1 --no-parity > gpurun_out/${S}_bench_b1.json 2> gpurun_out/${S}_b1.err
timeout 900 python bench.py --steps 10 --workload c2 --no-parity > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_c2.err
for f in n1 b1 c2; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
print("   ", {k: (v["ms_per_step"], v.get("frac_of_hbm_peak")) for k, v in d["roofline"]["families"].items()})
PY
done
tail -n 2 gpurun_out/${S}_n1.err
