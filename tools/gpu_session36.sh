#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q --no-header 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -8 | cut -c1-250
for pk in 1 0; do
AG3D_S2C_PACK=$pk timeout 200 python bench.py --steps 8 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pack', $pk, round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['families']['s2c_mask']['ms_per_step'])"
done
