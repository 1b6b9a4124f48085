#!/bin/bash
# 2-GPU sanity of the final bench.py (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n2_final.json 2> gpurun_out/r2_n2_final.err; echo "n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_n2_final.json').read().strip().splitlines()[-1])
    print('n2', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('params_in_sync'), d['clocks'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_n2_final.err').read()[-2000:])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_n2_ref.json 2> gpurun_out/r2_n2_ref.err; echo "n2 ref rc=$?"; tail -c 400 gpurun_out/r2_n2_ref.json
