#!/bin/bash
# One GPU visit (gpurun -- 'bash tools/gpu_visit.sh <what> [args]'); everything lands in gpurun_out/ with the given tag.
#   tests                      full `pytest -m gpu` + smoke
#   bench  <tag> [bench args]  bench.py on one GPU
#   multi  <N> <tag> [args]    bench.py under torchrun on N GPUs (+ the 2-GPU pytest when N >= 2 and tag ends in "t")
#   profile <tag>              ncu launch list of one eager denoise step (view batch 36) + ncu --set full of the attention
#                              kernel at B = 72 rows + ncu --set full of the persistent GEMM (GEGLU 320 -> 2560)
#                              + warm launch list of two eval renders
set -x
mkdir -p gpurun_out
what=$1; shift
case $what in
  tests)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/gputest.log
    python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 ;;
  bench)
    tag=$1; shift
    python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
    tail -3 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json ;;
  multi)
    N=$1; tag=$2; shift 2
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N "$@" > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
    tail -3 gpurun_out/${tag}_bench_n$N.err; cut -c1-400 gpurun_out/${tag}_bench_n$N.json ;;
  profile)
    tag=$1
    GCB_PROFILE_VB=36 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" \
      --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 1 > gpurun_out/${tag}_step.log 2>&1
    tail -1 gpurun_out/${tag}_step.log
    GCB_PROFILE_BQ=72 timeout 900 ncu --set full --import-source on --clock-control none -k regex:attn_tc_kernel -s 1 -c 1 \
      -o gpurun_out/${tag}_attn_b72 -f python tools/profile_attn.py 3 > gpurun_out/${tag}_attn.log 2>&1
    ncu -i gpurun_out/${tag}_attn_b72.ncu-rep --page details > gpurun_out/${tag}_attn_b72_ncu.txt 2>&1
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 \
      -o gpurun_out/${tag}_gemm_geglu320 -f python tools/profile_gemm.py > gpurun_out/${tag}_gemm.log 2>&1
    ncu -i gpurun_out/${tag}_gemm_geglu320.ncu-rep --page details > gpurun_out/${tag}_gemm_geglu320_ncu.txt 2>&1
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --cache-control none -c 45 --csv --log-file gpurun_out/${tag}_raster_launches.csv python tools/time_raster.py 1000000 2 \
      > gpurun_out/${tag}_raster.log 2>&1 ;;
esac
