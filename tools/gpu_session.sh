#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/s12_pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.log 2>&1
timeout 900 python bench.py --steps 8 > gpurun_out/s12_bench_forward_b8.json 2> gpurun_out/s12_bench_forward_b8.err
timeout 600 python bench.py --steps 8 --batch 1 --no-parity > gpurun_out/s12_bench_forward_b1.json 2> gpurun_out/s12_b1.err
timeout 600 python bench.py --steps 8 --workload c2 --no-parity > gpurun_out/s12_bench_c2.json 2> gpurun_out/s12_c2.err
timeout 900 python bench.py --steps 3 --warmup 1 --workload clickloop > gpurun_out/s12_bench_clickloop.json 2> gpurun_out/s12_clickloop.err
timeout 900 python bench.py --steps 5 --workload train > gpurun_out/s12_bench_train.json 2> gpurun_out/s12_train.err
tail -n 5 gpurun_out/s12_pytest_gpu.log; tail -n 3 gpurun_out/s12_smoke.log
for f in forward_b8 forward_b1 c2 clickloop train; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s12_bench_$f.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d}, d.get("e2e"), d.get("parity"))
    r = d.get("roofline") or {}
    for k, v in (r.get("families") or {}).items():
        print("   ", k, v.get("ms_per_step"), v.get("frac_of_hbm_peak"), v.get("binding"), v.get("frac_of_binding_bound"))
    print("   cfg", {k: v for k, v in d.get("config", {}).items() if k != "workload"})
except Exception as e:
    print("no json:", e)
PY
done
tail -n 5 gpurun_out/s12_*.err
