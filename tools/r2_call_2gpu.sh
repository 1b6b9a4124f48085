#!/bin/bash
# 2-GPU data-parallel check: overlapped bucketed all-reduce vs the single all-reduce, params_in_sync
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline; }
UNFLOW_OVERLAP_ALLREDUCE=1 run 29511 > gpurun_out/r2_n2_overlap.json 2> gpurun_out/r2_n2_overlap.err
UNFLOW_OVERLAP_ALLREDUCE=0 run 29512 > gpurun_out/r2_n2_single.json 2> gpurun_out/r2_n2_single.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_n1_same_box.json 2> gpurun_out/r2_n1_same_box.err
python - <<'PY'
import json
for f in ('r2_n2_overlap','r2_n2_single','r2_n1_same_box'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'in_sync', d.get('params_in_sync'), 'loss', d['final_loss'])
    except Exception as e:
        print(f,'FAILED',e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
