#!/bin/bash
mkdir -p gpurun_out
python tools/profile_vae.py 4
python tools/profile_vae.py 8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv --log-file gpurun_out/r2o_vae_launches.csv python tools/profile_vae.py 4 > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log
