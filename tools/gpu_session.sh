#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s41
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${S}_smoke.log 2>&1; tail -n 1 gpurun_out/${S}_smoke.log
timeout 900 python bench.py --steps 12 > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --steps 12 --no-parity --batch 1 > gpurun_out/${S}_bench_b1.json 2> gpurun_out/${S}_b1.err
timeout 900 python bench.py --steps 12 --no-parity --workload c2 > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_c2.err
for f in n1 b1 c2; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d.get("parity") and d["parity"]["mask_logits_rel_err_per_layer"])
PY
done
tail -n 3 gpurun_out/${S}_n1.err
