#!/bin/bash
mkdir -p gpurun_out
python tools/ab_attn_libs.py gaussctrl_b200/libgcb_attn_two.so gaussctrl_b200/libgcb_attn_two_st4.so 2>&1 | tee gpurun_out/r2k_ab_attn.txt
