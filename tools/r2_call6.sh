#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python tools/tc_conv_check.py --wgrad-only > gpurun_out/tc_check_v4.jsonl 2> gpurun_out/tc_check_v4.err
echo "check rc=$?"; python - <<'PY'
import json,sys
ok=True; seen=0
for ln in open('gpurun_out/tc_check_v4.jsonl'):
    d=json.loads(ln)
    if 'case' in d and d['mode']=='wgrad':
        seen+=1
        print(d['case'], 'err %.2e'%d['err'], d.get('us'), d.get('tflops_fp32_equiv'), d['finite'])
        if not (d['err']<1e-4): ok=False
open('gpurun_out/wgrad_ok','w').write('1' if (ok and seen>=6) else '0')
PY
tail -5 gpurun_out/tc_check_v4.err
export UNFLOW_TC_WGRAD=$(cat gpurun_out/wgrad_ok); echo "UNFLOW_TC_WGRAD=$UNFLOW_TC_WGRAD"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest8.log 2>&1
echo "pytest rc=$?"; grep -E "full-size|CSS 384|passed|failed|FAILED|Error" gpurun_out/r2_pytest8.log | tail -30
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v5.json 2> gpurun_out/r2_tc_v5.err; tail -3 gpurun_out/r2_tc_v5.err
python - <<'PY'
import json
for f in ('r2_tc_v5',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
    except Exception as e:
        print(f,'FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v5.md 2> gpurun_out/r2_kernel_table_v5.err; head -45 gpurun_out/r2_kernel_table_v5.md
