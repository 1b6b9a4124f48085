#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_default5.json 2> gpurun_out/r2_bench_default5.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_default5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default5.json').read().strip().splitlines()[-1])
print('final', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['rooflines_other'][0]['frac'], d['gpu_launches'], d.get('fp32_exact'), d['cpu_baseline']['value'])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --also-fp32 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('short', d['ms_per_step'], d['clocks'])"
