#!/bin/bash
# 8-GPU data-parallel check: overlapped bucketed all-reduce vs the single all-reduce, params_in_sync
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0; }
UNFLOW_OVERLAP_ALLREDUCE=1 run 29511 > gpurun_out/r2_n8_overlap.json 2> gpurun_out/r2_n8_overlap.err
UNFLOW_OVERLAP_ALLREDUCE=0 run 29512 > gpurun_out/r2_n8_single.json 2> gpurun_out/r2_n8_single.err
python - <<'PY'
import json
for f in ('r2_n8_overlap','r2_n8_single'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'in_sync', d.get('params_in_sync'), 'loss', d['final_loss'])
    except Exception as e:
        print(f,'FAILED',e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
