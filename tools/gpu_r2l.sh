#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_attn_libs.py gaussctrl_b200/libgaussctrl_b200.so gaussctrl_b200/libgcb_attn_pp.so 2>&1 | tee gpurun_out/r2l_ab_attn.txt
