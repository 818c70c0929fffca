#!/bin/bash
# round 2, visit b: new binning (parity) + raster timing + per-kernel launch list of one view
set -x
mkdir -p gpurun_out
python -m pytest tests/test_raster_gpu.py tests/test_z_fullsize_properties_gpu.py tests/test_z_fullsize_elementwise_gpu.py tests/test_pipeline_gpu.py tests/test_finetune_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
python tools/time_raster.py 1000000 40 > gpurun_out/r2b_raster_time.json 2>gpurun_out/r2b_raster_time.err
cat gpurun_out/r2b_raster_time.json; tail -3 gpurun_out/r2b_raster_time.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_raster_launches.csv python tools/time_raster.py 1000000 2 > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log
