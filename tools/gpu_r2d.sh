#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"radix_pass_kernel|emit_isects_kernel|rasterize_fwd_kernel|gather_scan_kernel" --launch-skip 8 --launch-count 8 -o gpurun_out/r2d_raster_full -f python tools/time_raster.py 1000000 2 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
ls -la gpurun_out/
