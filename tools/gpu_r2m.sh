#!/bin/bash
# 8-GPU visit: strong scaling on cfg2 (V=40) at N=8 and N=4
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 2 --no-extras > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
tail -3 gpurun_out/r2m_bench_n$N.err
cat gpurun_out/r2m_bench_n$N.json | cut -c1-400
done
