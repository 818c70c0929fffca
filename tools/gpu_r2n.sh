#!/bin/bash
# 2-GPU visit: overlapped (double-buffered) reference pass: parity + N=2 bench; plus single-GPU denoiser tests
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-extras > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
tail -5 gpurun_out/r2n_bench_n2.err
cut -c1-300 gpurun_out/r2n_bench_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
tail -3 gpurun_out/r2n_bench_n1.err
cut -c1-300 gpurun_out/r2n_bench_n1.json
