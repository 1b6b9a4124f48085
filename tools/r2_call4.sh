#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest4.log 2>&1
echo "pytest rc=$?"; grep -E "full-size|CSS|passed|failed|FAILED|Error" gpurun_out/r2_pytest4.log | tail -30
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v1.json 2> gpurun_out/r2_tc_v1.err; tail -c 1500 gpurun_out/r2_tc_v1.json; tail -3 gpurun_out/r2_tc_v1.err
UNFLOW_TC_CONV=0 $B > gpurun_out/r2_tc_off.json 2> gpurun_out/r2_tc_off.err
python - <<'PY'
import json
for f in ('r2_tc_v1','r2_tc_off'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
    except Exception as e:
        print(f,'FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v1.md 2> gpurun_out/r2_kernel_table_v1.err; head -50 gpurun_out/r2_kernel_table_v1.md
