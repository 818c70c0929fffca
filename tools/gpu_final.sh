#!/bin/bash
# Round snapshot on one B200 (~5 GPU-minutes): whole GPU suite, smoke, default bench, attention A/B of the unmeasured
# 4-piece P-store split, ncu launch list of one denoise step at the bench's launch shapes.
# Before calling: bash tools/build_attn_variants.sh "st4=-DGCB_ATTN_SPLIT_ST=4"   (variant .so files travel with gpurun)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 8 gpurun_out/all_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "== bench exit $?"; head -c 600 gpurun_out/bench_full.json; echo
if ls gaussctrl_b200/libgcb_attn_*.so > /dev/null 2>&1; then
  timeout 120 python tools/ab_attn_libs.py gaussctrl_b200/libgaussctrl_b200.so gaussctrl_b200/libgcb_attn_*.so > gpurun_out/ab_attn.txt 2>&1; cat gpurun_out/ab_attn.txt
fi
GCB_PROFILE_VB=36 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py 1 > gpurun_out/profile_step.log 2>&1; echo "== launch list exit $?"
