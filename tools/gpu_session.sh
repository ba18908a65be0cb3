#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s52
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${S}_smoke.log 2>&1; tail -n 1 gpurun_out/${S}_smoke.log
timeout 900 python bench.py > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${S}_bench_reference.json 2> gpurun_out/${S}_ref.err
timeout 900 python bench.py --steps 10 --workload c2 --no-parity > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_c2.err
for f in n1 reference c2; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d.get("parity") and d["parity"]["mask_logits_rel_err_per_layer"])
PY
done
tail -n 2 gpurun_out/${S}_n1.err
