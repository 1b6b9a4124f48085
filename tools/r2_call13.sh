#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/tc_conv_check.py --ab-tmem > gpurun_out/tc_ab_tmem.jsonl 2> gpurun_out/tc_ab_tmem.err
echo "ab rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/tc_ab_tmem.jsonl'):
    d=json.loads(ln)
    if 'phase' in d: print('----', d['phase'])
    if 'case' in d: print('%-28s err %.2e %s us=%s us_wgrad=%s tf=%s'%(d['case'], d['err'], d.get('err_wgrad',''), d.get('us'), d.get('us_wgrad'), d.get('tflops_fp32_equiv')))
PY
tail -5 gpurun_out/tc_ab_tmem.err
