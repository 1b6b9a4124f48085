#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python tools/tc_conv_check.py > gpurun_out/tc_check_v2.jsonl 2> gpurun_out/tc_check_v2.err
echo "check rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/tc_check_v2.jsonl'):
    d=json.loads(ln)
    if 'case' in d: print(d['case'], 'err %.2e'%d['err'], d.get('us'), d.get('tflops_fp32_equiv'), 'cudnn3x', d.get('err_cudnn_3xtf32'), d.get('us_cudnn_3xtf32_with_operand_passes'), 'fp32', d.get('err_cudnn_fp32'), d['finite'], d['slack_untouched'])
PY
tail -5 gpurun_out/tc_check_v2.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest5.log 2>&1
echo "pytest rc=$?"; grep -E "full-size|CSS 384|passed|failed|FAILED|Error" gpurun_out/r2_pytest5.log | tail -30
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v2.json 2> gpurun_out/r2_tc_v2.err; tail -3 gpurun_out/r2_tc_v2.err
python - <<'PY'
import json
for f in ('r2_tc_v2',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
    except Exception as e:
        print(f,'FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v2.md 2> gpurun_out/r2_kernel_table_v2.err; head -60 gpurun_out/r2_kernel_table_v2.md
