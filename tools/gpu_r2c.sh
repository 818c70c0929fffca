#!/bin/bash
# warm-cache per-kernel timings of the rasteriser (ncu replays with --cache-control none)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 45 --csv --log-file gpurun_out/r2c_raster_launches_warm.csv python tools/time_raster.py 1000000 2 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
