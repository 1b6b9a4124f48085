#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/probes/mma_probe > gpurun_out/mma_probe.jsonl 2> gpurun_out/mma_probe.err; echo "probe rc=$?"; cat gpurun_out/mma_probe.jsonl | cut -c1-220; tail -3 gpurun_out/mma_probe.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2_pytest27.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest27.log | tail -12
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_tc_v17.json 2> gpurun_out/r2_tc_v17.err; tail -3 gpurun_out/r2_tc_v17.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_tc_v17.json').read().strip().splitlines()[-1])
    print('v17', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])
    print(json.dumps(d['roofline'])[:300]); print(json.dumps(d['rooflines_other'][0])[:300])
except Exception as e: print('FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_v17.md 2> gpurun_out/r2_kernel_table_v17.err; head -24 gpurun_out/r2_kernel_table_v17.md
timeout 200 python tools/bench_ops.py > gpurun_out/r2_bench_ops.jsonl 2> gpurun_out/r2_bench_ops.err; grep -i "forward_warp\|image_warp" gpurun_out/r2_bench_ops.jsonl | cut -c1-300
