#!/bin/bash
# r1n: ncu launch list of one eager denoise step with the final kernels (reference pass CFG batch 8 + view batch CFG batch 72)
mkdir -p gpurun_out
GCB_PROFILE_VB=36 timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv \
    --log-file gpurun_out/launches_r1n.csv python tools/profile_step.py 1 > gpurun_out/profile_step_r1n.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_r1n.csv; tail -n 2 gpurun_out/profile_step_r1n.log
