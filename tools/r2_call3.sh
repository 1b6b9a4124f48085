#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/tc_conv_check.py --quick > gpurun_out/tc_check_quick.jsonl 2> gpurun_out/tc_check_quick.err
echo "quick rc=$?"; cat gpurun_out/tc_check_quick.jsonl; tail -5 gpurun_out/tc_check_quick.err
timeout 400 python tools/tc_conv_check.py > gpurun_out/tc_check.jsonl 2> gpurun_out/tc_check.err
echo "full rc=$?"; tail -20 gpurun_out/tc_check.jsonl; tail -5 gpurun_out/tc_check.err
