#!/bin/bash
# ncu passes of the profiling recipe (B200_PROFILING.md): launch list of one denoise step, then --set full of the top kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py 1 > gpurun_out/profile_step.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 2 \
    -o gpurun_out/prof_attn_tc -f python tools/profile_step.py 1 > gpurun_out/profile_attn.log 2>&1
echo "attn full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 \
    -o gpurun_out/prof_gemm_tc -f python tools/profile_step.py 1 > gpurun_out/profile_gemm.log 2>&1
echo "gemm full exit $?"; ls -la gpurun_out/*.ncu-rep
