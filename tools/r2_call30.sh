#!/bin/bash
# round-2 evidence pass: accumulation-chunk sweep, default bench lines (C and CSS), reference arm, ncu launch list
# of the bench command, ncu --set full of the pair kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/tc_conv_check.py --chunk-test > gpurun_out/tc_chunk_v2.jsonl 2> gpurun_out/tc_chunk_v2.err
python - <<'PY'
import json
for ln in open('gpurun_out/tc_chunk_v2.jsonl'):
    d=json.loads(ln)
    if 'case' in d: print('%-28s err %.2e us=%s tf=%s'%(d['case'], d['err'], d.get('us') or d.get('us_wgrad'), d.get('tflops_fp32_equiv')))
    else: print(d)
PY
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench default rc=$?"; tail -c 1500 gpurun_out/r2_bench_default.json | cut -c1-1500
timeout 400 python bench.py --spec CSS --batch 2 --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_bench_css.json 2> gpurun_out/r2_bench_css.err; echo "bench css rc=$?"; tail -c 600 gpurun_out/r2_bench_css.json; tail -3 gpurun_out/r2_bench_css.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "bench ref rc=$?"; tail -c 700 gpurun_out/r2_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_launches_step.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2_launches_step.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -o gpurun_out/r2_prof_tc_conv_pair -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_conv_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/r2_prof_tc_wgrad_pair -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_wgrad_pair.log 2>&1
ls -la gpurun_out/*.ncu-rep
