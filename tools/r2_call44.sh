#!/bin/bash
# final evidence pass of round 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest_final.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_final.log | tail -12
timeout 600 python bench.py > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err; echo "bench default rc=$?"
timeout 400 python bench.py --spec CSS --batch 2 --steps 10 --warmup 3 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_bench_css2.json 2> gpurun_out/r2_bench_css2.err; echo "bench css rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_default2','r2_bench_css2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['rooflines_other'][0]['frac'], d.get('fp32_exact'))
    except Exception as e: print(f,'FAILED',e)
PY
timeout 300 python tools/kernel_time_table.py > gpurun_out/r2_kernel_table_final.md 2> gpurun_out/r2_kernel_table_final.err; head -34 gpurun_out/r2_kernel_table_final.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_step2.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --also-fp32 0 > gpurun_out/r2_launches_step2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2_launches_step2.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -o gpurun_out/r2_prof_tc_conv_final -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_conv_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:narrow_fwd_tma -c 1 -o gpurun_out/r2_prof_narrow_fwd_tma -f python tools/bench_ops.py > gpurun_out/r2_prof_narrow.log 2>&1
timeout 200 python tools/bench_ops.py > gpurun_out/r2_bench_ops2.jsonl 2> gpurun_out/r2_bench_ops2.err; grep -i "narrow\|forward_warp" gpurun_out/r2_bench_ops2.jsonl | cut -c1-260
ls -la gpurun_out/*final*.ncu-rep gpurun_out/r2_prof_narrow_fwd_tma.ncu-rep 2>&1 | tail -3
