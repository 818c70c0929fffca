#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --views 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench_small exit $?"; tail -c 3000 gpurun_out/bench_small.json; tail -n 15 gpurun_out/bench_small.err
