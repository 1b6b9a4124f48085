#!/bin/bash
# same-box A/B of two builds of the weight-gradient kernel (512 vs 640 threads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 640 768 640 768; do
  cp unflow_b200/libunflow_$v.so unflow_b200/libunflow.so
  timeout 200 python tools/tc_conv_check.py --wgrad-only > gpurun_out/tc_wgrad_ab_$v.jsonl 2> gpurun_out/tc_wgrad_ab_$v.err
  python - <<PY
import json
out=[]
for ln in open('gpurun_out/tc_wgrad_ab_$v.jsonl'):
    d=json.loads(ln)
    if 'case' in d and (d.get('us') or d.get('us_wgrad')): out.append('%s=%s'%(d['case'].replace('wgrad ',''), d.get('us') or d.get('us_wgrad')))
print('$v', ' '.join(out))
PY
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_ab_$v.json 2> gpurun_out/r2_ab_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_ab_$v.json').read().strip().splitlines()[-1]); print('$v step', d['ms_per_step'], d['e2e']['ms_per_step'], d['rooflines_other'][0]['frac'], d['clocks'])"
done
cp unflow_b200/libunflow_640.so unflow_b200/libunflow.so
