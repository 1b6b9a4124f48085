#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/tc_conv_check.py --roles > gpurun_out/tc_roles_v6.jsonl 2> gpurun_out/tc_roles_v6.err
echo "roles rc=$?"; cat gpurun_out/tc_roles_v6.jsonl | cut -c1-520; tail -5 gpurun_out/tc_roles_v6.err
