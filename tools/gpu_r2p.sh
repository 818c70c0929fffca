#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2p_gputest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
