#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench_full exit $?"; tail -c 3500 gpurun_out/bench_full.json; tail -n 8 gpurun_out/bench_full.err
