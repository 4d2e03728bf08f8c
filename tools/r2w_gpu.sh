#!/bin/bash
# gpurun call W: intermediate carve-outs for the small kernels only (the sweep keeps its L1)
mkdir -p gpurun_out
for v in "0:0,256" "86:0,0" "86:0,256" "100:0,0" "100:0,256" "75:0,0" "90:0,0"; do
  IFS=: read c w <<< "$v"
  echo "## DVT_FIT_CARVEOUT=$c DVT_FIT_X3_WIDE_MIN_N=$w"
  DVT_FIT_CARVEOUT=$c DVT_FIT_X3_WIDE_MIN_N=$w timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done > gpurun_out/r2w_carveout.txt 2>&1
cat gpurun_out/r2w_carveout.txt
