#!/bin/bash
# 2-GPU visit: peer-memory K/V exchange parity + strong-scaling bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-extras > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -5 gpurun_out/r2h_bench_n2.err
cat gpurun_out/r2h_bench_n2.json
