#!/bin/bash
# r1k: attention slot-lag experiment (GCB_ATTN_LAG = 0 default / 1 / 2 as separate libraries): timing + parity
mkdir -p gpurun_out
for v in default lag1 lag2; do
  if [ $v = default ]; then unset GCB_LIB_PATH; else export GCB_LIB_PATH=$PWD/gaussctrl_b200/libgcb_attn_$v.so; fi
  timeout 200 python tools/time_attn.py > gpurun_out/time_attn_$v.txt 2>&1; echo "== time_attn $v exit $?"; grep tcgen05 gpurun_out/time_attn_$v.txt | grep -v ones
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --timeout 200 --tb=short -k "attention" > gpurun_out/attn_tests_$v.log 2>&1
  echo "== attention tests $v exit $?"; tail -n 2 gpurun_out/attn_tests_$v.log | cut -c1-200
done
unset GCB_LIB_PATH
