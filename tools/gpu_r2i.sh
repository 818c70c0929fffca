#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_processor_gpu.py tests/test_z_fullsize_properties_gpu.py tests/test_z_fullsize_elementwise_gpu.py -x -q 2>&1 | tail -8
python tools/ab_attn_libs.py gaussctrl_b200/libgcb_attn_two.so gaussctrl_b200/libgaussctrl_b200.so 2>&1 | tee gpurun_out/r2i_ab_attn_slots.txt
