#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_pipeline_gpu.py tests/test_denoiser_gpu.py -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -5 gpurun_out/r2g_bench.err
cat gpurun_out/r2g_bench.json
