#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest34.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest34.log | tail -12
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v22.json 2> gpurun_out/r2_tc_v22.err; tail -3 gpurun_out/r2_tc_v22.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_tc_v22.json').read().strip().splitlines()[-1])
    print('v22', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
except Exception as e: print('FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v22.md 2> gpurun_out/r2_kernel_table_v22.err; head -30 gpurun_out/r2_kernel_table_v22.md
