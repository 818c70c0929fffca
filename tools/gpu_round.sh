#!/bin/bash
# One GPU visit: kernel parity tests per file (isolated processes, bounded by timeout), logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --timeout 300 --tb=short 2>&1 > gpurun_out/$name.log
  echo "== $name: exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$name.log | cut -c1-220 | tail -n 40
done
