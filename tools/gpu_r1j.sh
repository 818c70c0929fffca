#!/bin/bash
# r1j: attention MMA issue reorder (both QK^T ahead of the PVs) A/B vs the r1h library; GEMM tile knobs; suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short -k "attention or batch_invariance" > gpurun_out/attn_tests.log 2>&1
echo "== attention tests exit $?"; tail -n 5 gpurun_out/attn_tests.log | cut -c1-300
timeout 200 python tools/time_attn.py > gpurun_out/time_attn_new.txt 2>&1; echo "== time_attn new exit $?"; cat gpurun_out/time_attn_new.txt | grep -v ones
GCB_LIB_PATH=$PWD/gaussctrl_b200/libgcb_r1h_baseline.so timeout 200 python tools/time_attn.py > gpurun_out/time_attn_r1h.txt 2>&1; echo "== time_attn r1h exit $?"; cat gpurun_out/time_attn_r1h.txt | grep -v ones
GCB_GEMM_TABLE_OUT=gpurun_out/gemm_table_geglu128.json GCB_GEGLU_BN=128 timeout 300 python tools/time_gemms.py > gpurun_out/gemm_table_geglu128.txt 2>&1; echo "== geglu128 exit $?"; grep " 1 2 " gpurun_out/gemm_table_geglu128.txt | head -8 | cut -c1-130; tail -n 1 gpurun_out/gemm_table_geglu128.txt
GCB_GEMM_TABLE_OUT=gpurun_out/gemm_table_maxbn128.json GCB_GEMM_MAX_BN=128 timeout 300 python tools/time_gemms.py > gpurun_out/gemm_table_maxbn128.txt 2>&1; echo "== maxbn128 exit $?"; tail -n 1 gpurun_out/gemm_table_maxbn128.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 4 gpurun_out/all_tests.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; echo "== bench exit $?"; head -c 300 gpurun_out/bench_r1j.json; echo
python - <<'PY'
import json
b = json.load(open("gpurun_out/bench_r1j.json"))
print("e2e", b["e2e"]["value"], "roofline", b["roofline"]["achieved"], b["roofline"]["frac"], "breakdown", b["extra"]["breakdown"])
PY
