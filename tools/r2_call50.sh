#!/bin/bash
# last evidence pass: final tree (20-warp weight-gradient kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest50.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest50.log | tail -8
timeout 600 python bench.py > gpurun_out/r2_bench_default4.json 2> gpurun_out/r2_bench_default4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default4.json').read().strip().splitlines()[-1])
print('final', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['rooflines_other'][0]['frac'], d['gpu_launches'])
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_final2.md 2> gpurun_out/r2_kernel_table_final2.err; head -12 gpurun_out/r2_kernel_table_final2.md
timeout 100 python tools/tc_conv_check.py --roles > gpurun_out/tc_roles_v7.jsonl 2> gpurun_out/tc_roles_v7.err; grep wgrad gpurun_out/tc_roles_v7.jsonl | cut -c1-420
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/r2_prof_tc_wgrad_final -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_wgrad_final.log 2>&1; ls -la gpurun_out/r2_prof_tc_wgrad_final.ncu-rep
