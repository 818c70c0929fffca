#!/bin/bash
for kb in 110 150 200; do
  GCB_GEMM_SMEM_KB=$kb timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('smem_kb=$kb', round(d['value'],3), 'views/s', round(d['ms_per_step']), 'ms')
"
done
