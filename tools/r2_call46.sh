#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest46.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest46.log | tail -8
timeout 600 python bench.py > gpurun_out/r2_bench_default3.json 2> gpurun_out/r2_bench_default3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default3.json').read().strip().splitlines()[-1])
print('final', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['rooflines_other'][0]['frac'], d['gpu_launches'])
PY
