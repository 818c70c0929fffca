#!/bin/bash
# round snapshot: full GPU test suite, smoke, default bench, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_full.json
bash tools/gpu_profile.sh 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv --log-file gpurun_out/raster_launches.csv python tools/profile_raster.py 1000000 2 > /dev/null 2>&1; echo "raster ncu exit $?"
