#!/bin/bash
# Round-end verification on one B200 (run through gpurun): A/B of the two staging loaders of the
# narrow-conv kernels, then the GPU test suite, smoke, the bench lines and the kernel time table
# under the faster loader.  Everything is written to gpurun_out/f_*.
mkdir -p gpurun_out
UNFLOW_NARROW_LOADER=1 python -m pytest tests/test_gpu_conv3x.py -x -q -k narrow > gpurun_out/f_narrow_async.log 2>&1
rc=$?
tail -2 gpurun_out/f_narrow_async.log
L=0
if [ $rc -eq 0 ]; then
  UNFLOW_NARROW_LOADER=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_l1.json 2> gpurun_out/f_bench_l1.err
  UNFLOW_NARROW_LOADER=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_l0.json 2> gpurun_out/f_bench_l0.err
  L=$(python - <<'PY'
import json
try:
    a = json.load(open('gpurun_out/f_bench_l1.json'))['ms_per_step']
    b = json.load(open('gpurun_out/f_bench_l0.json'))['ms_per_step']
    print(1 if a < 0.997 * b else 0)
except Exception:
    print(0)
PY
)
fi
export UNFLOW_NARROW_LOADER=$L
echo "loader=$L" | tee gpurun_out/f_choice.txt
python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest.log 2>&1; tail -2 gpurun_out/f_pytest.log
python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cut -c1-220 gpurun_out/f_bench.json
python tools/kernel_time_table.py > gpurun_out/f_kernel_table.md 2> gpurun_out/f_kernel_table.err; head -12 gpurun_out/f_kernel_table.md
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; cut -c1-200 gpurun_out/f_ref.json
