#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest14.log 2>&1
echo "pytest rc=$?"; grep -E "full-size|CSS 384|passed|failed|FAILED|Error" gpurun_out/r2_pytest14.log | tail -30
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v9.json 2> gpurun_out/r2_tc_v9.err; tail -3 gpurun_out/r2_tc_v9.err
python - <<'PY'
import json
for f in ('r2_tc_v9',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
        print(json.dumps(d['roofline'])); 
        for r in d['rooflines_other'][:3]: print(json.dumps(r))
    except Exception as e:
        print(f,'FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v9.md 2> gpurun_out/r2_kernel_table_v9.err; head -32 gpurun_out/r2_kernel_table_v9.md
sed -i 's/r2_prof_tc_conv/r2_prof_tc_conv_at/g; s/r2_prof_tc_wgrad/r2_prof_tc_wgrad_at/g' tools/r2_ncu_tc.sh
bash tools/r2_ncu_tc.sh
