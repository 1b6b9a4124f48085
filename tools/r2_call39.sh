#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/tc_conv_check.py --pair-test > gpurun_out/tc_pair_v5.jsonl 2> gpurun_out/tc_pair_v5.err
echo "pair rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/tc_pair_v5.jsonl'):
    d=json.loads(ln)
    if 'case' in d: print('%-50s err %.2e us=%s tf=%s'%(d['case'], d['err'], d.get('us') or d.get('us_wgrad'), d.get('tflops_fp32_equiv')))
    else: print(d)
PY
tail -3 gpurun_out/tc_pair_v5.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest39.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest39.log | tail -12
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v26.json 2> gpurun_out/r2_tc_v26.err; tail -3 gpurun_out/r2_tc_v26.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_tc_v26.json').read().strip().splitlines()[-1])
    print('v26', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
except Exception as e: print('FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v26.md 2> gpurun_out/r2_kernel_table_v26.err; head -16 gpurun_out/r2_kernel_table_v26.md
