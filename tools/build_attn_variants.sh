#!/bin/bash
# Build A/B variants of the attention kernel as separate shared libraries next to the default one:
#   libgcb_attn_<name>.so for name=flags pairs given as arguments, e.g.  ld=-DGCB_ATTN_SPLIT_LD=1
set -e
cd "$(dirname "$0")/../gaussctrl_b200"
OBJS=$(ls _build/*.o | grep -v attn_tc.o | grep -v variant)
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       $flags -c csrc/attn_tc.cu -o _build/attn_tc_variant_$name.o
  nvcc -shared -o libgcb_attn_$name.so $OBJS _build/attn_tc_variant_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
done
ls -la libgcb_attn_*.so
