#!/bin/bash
# gpurun call Z1: persisting L2 window over the chain's arena
mkdir -p gpurun_out
DVT_FIT_DEBUG_GRAPH=1 DVT_FIT_L2_PERSIST_MB=64 timeout 600 python tools/fit_breakdown.py --iters 200 --graphs-only --graph-steps 20 2>&1 | grep "L2 window" | head -2
for mb in 0 24 48 64 96; do
  echo "## DVT_FIT_L2_PERSIST_MB=$mb"
  DVT_FIT_L2_PERSIST_MB=$mb timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done > gpurun_out/r2z1_l2.txt 2>&1
cat gpurun_out/r2z1_l2.txt
DVT_FIT_L2_PERSIST_MB=64 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2z1_fit_timeline_l2.csv 2>&1 | tail -1
DVT_FIT_L2_PERSIST_MB=64 timeout 600 python -m pytest tests/test_fit_gpu.py -x -q 2>&1 | tail -2
