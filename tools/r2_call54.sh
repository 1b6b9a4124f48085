#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_tc_conv.py -m gpu -q --timeout 300 2>&1 | tail -2
