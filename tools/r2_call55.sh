#!/bin/bash
cd "$(dirname "$0")/.."
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also-fp32 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ok', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d['clocks'])"
