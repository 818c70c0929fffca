#!/bin/bash
mkdir -p gpurun_out
python tools/ab_attn_libs.py gaussctrl_b200/libgcb_attn_two.so gaussctrl_b200/libgaussctrl_b200.so gaussctrl_b200/libgcb_attn_bn64two.so gaussctrl_b200/libgcb_attn_order1.so gaussctrl_b200/libgcb_attn_two_order1.so 2>&1 | tee gpurun_out/r2j_ab_attn.txt
