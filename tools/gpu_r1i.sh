#!/bin/bash
# r1i: pipelined GEMM epilogue (TMEM loads in flight under the conversion, lean GEGLU / SiLU math): kernel parity,
# per-shape timing vs the r1h library, attention source-level capture, whole GPU suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_vae_gpu.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short > gpurun_out/kernel_tests.log 2>&1
echo "== kernel tests exit $?"; tail -n 12 gpurun_out/kernel_tests.log | cut -c1-300
timeout 300 python tools/time_gemms.py > gpurun_out/gemm_table.txt 2>&1; echo "== gemm table exit $?"; head -n 14 gpurun_out/gemm_table.txt | cut -c1-150; tail -n 1 gpurun_out/gemm_table.txt
GCB_GEMM_TABLE_OUT=gpurun_out/gemm_table_r1h_lib.json GCB_LIB_PATH=$PWD/gaussctrl_b200/libgcb_r1h_baseline.so timeout 300 python tools/time_gemms.py > gpurun_out/gemm_table_r1h_lib.txt 2>&1; echo "== baseline gemm table exit $?"; tail -n 1 gpurun_out/gemm_table_r1h_lib.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -o gpurun_out/attn_b24 -f python tools/profile_attn.py 3 > gpurun_out/attn_ncu.log 2>&1; echo "== attn ncu exit $?"; ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 6 gpurun_out/all_tests.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; echo "== bench exit $?"; head -c 400 gpurun_out/bench_r1i.json
