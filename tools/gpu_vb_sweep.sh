#!/bin/bash
for vb in 4 6 9 12 18; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --view-batch $vb 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('vb=$vb', round(d['value'],3), 'views/s', round(d['ms_per_step']), 'ms', round(d['roofline']['achieved']), 'TF attn')
"
done
