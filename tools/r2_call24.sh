#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2_pytest24.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest24.log | tail -12
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v14.json 2> gpurun_out/r2_tc_v14.err; tail -3 gpurun_out/r2_tc_v14.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_tc_v14.json').read().strip().splitlines()[-1])
    print('v14', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
    print(json.dumps(d['roofline'])[:300]); print(json.dumps(d['rooflines_other'][0])[:300])
except Exception as e: print('FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v14.md 2> gpurun_out/r2_kernel_table_v14.err; head -30 gpurun_out/r2_kernel_table_v14.md
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 2 -o gpurun_out/r2_prof_tc_conv64 -f python tools/tc_conv_check.py --profile2 > gpurun_out/r2_prof_tc_conv64.log 2>&1
ls -la gpurun_out/*.ncu-rep
