#!/bin/bash
# r1h: TMA-store GEMM epilogue - parity (both epilogues), per-shape A/B timing, then the whole GPU suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -k "conv or geglu or tma or batch_invariance" > gpurun_out/gemm_tests.log 2>&1
echo "== gemm tests exit $?"; tail -n 15 gpurun_out/gemm_tests.log | cut -c1-300
timeout 400 python tools/time_gemms.py > gpurun_out/gemm_table.txt 2>&1; echo "== gemm table exit $?"; tail -n 45 gpurun_out/gemm_table.txt | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 6 gpurun_out/all_tests.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo "== bench exit $?"; head -c 700 gpurun_out/bench_r1h.json
