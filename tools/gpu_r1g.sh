#!/bin/bash
# Round r1g GPU visit (tight budget): new parity tests first, then the default bench, the fine-tune launch list, and
# the full GPU suite with whatever time is left.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 420 python -m pytest tests/test_finetune_gpu.py tests/test_clip_gpu.py -q -m gpu -p no:cacheprovider --timeout 300 --tb=short > gpurun_out/new_tests.log 2>&1
echo "== new tests exit $?"; tail -n 25 gpurun_out/new_tests.log | cut -c1-300
timeout 200 python tools/profile_finetune.py > gpurun_out/finetune_timing.log 2>&1; echo "== finetune timing exit $?"; tail -n 4 gpurun_out/finetune_timing.log
timeout 600 python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "== bench exit $?"; tail -c 1500 gpurun_out/bench_r1g.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv --log-file gpurun_out/finetune_launches.csv python tools/profile_finetune.py > /dev/null 2>&1; echo "== finetune ncu exit $?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 6 gpurun_out/all_tests.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
