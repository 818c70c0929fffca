#!/bin/bash
# r1l: whole GPU suite on the final kernels, view-batch sweep, fine-tune launch list (forward + backward), smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --tb=short -x --deselect tests/test_multigpu_gpu.py > gpurun_out/all_tests.log 2>&1
echo "== all tests exit $?"; tail -n 4 gpurun_out/all_tests.log | cut -c1-300
for vb in 18 36; do
  timeout 300 python bench.py --view-batch $vb --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/bench_vb$vb.json 2> gpurun_out/bench_vb$vb.err
  echo "== vb $vb exit $?"; python -c "import json; b=json.load(open('gpurun_out/bench_vb$vb.json')); print('vb', $vb, 'views/s', b['value'], b['extra']['breakdown'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv --log-file gpurun_out/finetune_launches.csv python tools/profile_finetune.py > gpurun_out/finetune_ncu.log 2>&1; echo "== finetune ncu exit $?"; grep -c "gpu__time_duration" gpurun_out/finetune_launches.csv
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
