#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_raster_gpu.py -x -q 2>&1 | tail -3
for s in 1 2 3 4; do python tools/time_raster.py 1000000 40 $s 2>/dev/null | tee -a gpurun_out/r2f_raster_time.jsonl; done
