#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/tc_conv_check.py --wgrad-trunc > gpurun_out/tc_wgrad_trunc.jsonl 2> gpurun_out/tc_wgrad_trunc.err
python - <<'PY'
import json
for ln in open('gpurun_out/tc_wgrad_trunc.jsonl'):
    d=json.loads(ln)
    if 'case' in d: print('%-28s err %.2e us=%s tf=%s'%(d['case'], d['err'], d.get('us') or d.get('us_wgrad'), d.get('tflops_fp32_equiv')))
    else: print(d)
PY
tail -3 gpurun_out/tc_wgrad_trunc.err
