#!/bin/bash
# ncu --set full on one launch of the tcgen05 conv kernel (conv3_1 forward) and one of the weight-gradient kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -o gpurun_out/r2_prof_tc_conv -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/r2_prof_tc_wgrad -f python tools/tc_conv_check.py --profile > gpurun_out/r2_prof_tc_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
