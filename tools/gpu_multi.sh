#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider --timeout 600 --tb=short 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; tail -c 1500 gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
