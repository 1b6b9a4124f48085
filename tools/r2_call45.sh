#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/tc_conv_check.py --wgrad-only > gpurun_out/tc_wgrad_v10.jsonl 2> gpurun_out/tc_wgrad_v10.err
echo "wgrad rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/tc_wgrad_v10.jsonl'):
    d=json.loads(ln)
    if 'case' in d: print('%-40s err %.2e %s us=%s tf=%s'%(d['case'], d['err'], d.get('err_wgrad',''), d.get('us') or d.get('us_wgrad'), d.get('tflops_fp32_equiv')))
    else: print(d)
PY
tail -3 gpurun_out/tc_wgrad_v10.err
timeout 600 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_baseline_sizes.py tests/test_gpu_model.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest47.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest47.log | tail -8
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v28.json 2> gpurun_out/r2_tc_v28.err; tail -3 gpurun_out/r2_tc_v28.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_tc_v28.json').read().strip().splitlines()[-1])
    print('v28', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d['clocks'], d['rooflines_other'][0]['frac'])
except Exception as e: print('FAILED',e)
PY
