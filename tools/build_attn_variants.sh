#!/bin/bash
# Build A/B variants of the attention kernel (GCB_ATTN_LAG = 1, 2) as separate shared libraries next to the default one.
set -e
cd "$(dirname "$0")/../gaussctrl_b200"
OBJS=$(ls _build/*.o | grep -v attn_tc.o | grep -v variant)
for lag in 1 2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DGCB_ATTN_LAG=$lag -c csrc/attn_tc.cu -o _build/attn_tc_variant_lag$lag.o
  nvcc -shared -o libgcb_attn_lag$lag.so $OBJS _build/attn_tc_variant_lag$lag.o -gencode arch=compute_100a,code=sm_100a -lcudart
done
ls -la libgcb_attn_lag*.so
