#!/bin/bash
# end-of-round validation: build, smoke, the whole GPU suite, forward + train bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.log; cat gpurun_out/final_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 400 python bench.py --workload train --batch 8 --steps 3 --warmup 3 > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err
python - <<'PY'
import json
for f in ("final_bench","final_bench_train"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
