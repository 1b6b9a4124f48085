#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/tc_conv_check.py --ab-tmem --multicast > gpurun_out/tc_ab_mc.jsonl 2> gpurun_out/tc_ab_mc.err
echo "ab rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/tc_ab_mc.jsonl'):
    d=json.loads(ln)
    if 'phase' in d: print('----', d['phase'])
    if 'case' in d and d['mode']!='wgrad': print('%-28s err %.2e us=%s tf=%s'%(d['case'], d['err'], d.get('us'), d.get('tflops_fp32_equiv')))
PY
tail -5 gpurun_out/tc_ab_mc.err
