#!/bin/bash
echo "default (0x13):"; python tools/time_attn.py 2>&1 | sed -n 3,4p
for m in 0x00 0x11 0x33 0x37; do echo "mask $m:"; GCB_LIB_PATH=$PWD/gaussctrl_b200/libgcb_poly_$m.so python tools/time_attn.py 2>&1 | sed -n 3,4p; done
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tcgen05_attention" -p no:cacheprovider 2>&1 | tail -2
GCB_LIB_PATH=$PWD/gaussctrl_b200/libgcb_poly_0x37.so timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tcgen05_attention" -p no:cacheprovider 2>&1 | tail -2
