#!/bin/bash
# gpurun call O: off-path kernel priority + sweep PDL, timeline of the new default
mkdir -p gpurun_out
for prio in 0 1 2; do for spdl in 1 0; do
  echo "## DVT_FIT_OFFPATH_PRIO=$prio DVT_FIT_SWEEP_PDL=$spdl"
  DVT_FIT_OFFPATH_PRIO=$prio DVT_FIT_SWEEP_PDL=$spdl timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20;1:48,48:20;1:56,56:20;1:64,64:20' 2>&1 | grep -v "^+" | paste - -
done; done > gpurun_out/r2o_prio.txt 2>&1
cat gpurun_out/r2o_prio.txt
DVT_FIT_SWEEP_CTAS=48,48 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2o_fit_timeline_prio1.csv 2>&1 | tail -1
DVT_FIT_SWEEP_CTAS=48,48 DVT_FIT_SWEEP_PDL=0 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2o_fit_timeline_prio1_nopdl.csv 2>&1 | tail -1
timeout 900 python -m pytest tests/test_fit_gpu.py -x -q 2>&1 | tail -3
