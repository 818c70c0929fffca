#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_raster_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -5
for s in 1 2 4 6 8; do python tools/time_raster.py 1000000 40 $s 2>&1 | tail -1 | tee -a gpurun_out/r2q_raster_time.jsonl; done
