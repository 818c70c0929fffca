#!/bin/bash
for bn in 256 160 128; do
  GCB_GEMM_MAX_BN=$bn timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('max_bn=$bn', round(d['value'],3), 'views/s', round(d['ms_per_step']), 'ms')
"
done
