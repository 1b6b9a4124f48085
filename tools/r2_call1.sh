#!/bin/bash
# Round-2 first GPU call: full -m gpu suite (reference kernels un-gated, BASELINE-size parity),
# op-level timing against the reference kernels, A/B of the switches left unmeasured in round 1.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -s -x --timeout 600 > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 300 python tools/bench_reference_kernels.py > gpurun_out/r2_refkernels.jsonl 2> gpurun_out/r2_refkernels.err
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0"
$B > gpurun_out/r2_ab_base.json 2> gpurun_out/r2_ab_base.err
UNFLOW_CONV1_S2D=0 $B > gpurun_out/r2_ab_s2d0.json 2> gpurun_out/r2_ab_s2d0.err
UNFLOW_BWD_STREAMS=1 $B > gpurun_out/r2_ab_bwdstreams.json 2> gpurun_out/r2_ab_bwdstreams.err
UNFLOW_NARROW_LOADER=2 $B > gpurun_out/r2_ab_loader2.json 2> gpurun_out/r2_ab_loader2.err
$B --prefetch 1 > gpurun_out/r2_ab_prefetch.json 2> gpurun_out/r2_ab_prefetch.err
for f in base s2d0 bwdstreams loader2 prefetch; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_ab_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['e2e']['ms_per_step'])
except Exception as e:
    print(f,'FAILED',e)
PY
done
