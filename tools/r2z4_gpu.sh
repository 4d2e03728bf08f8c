#!/bin/bash
# gpurun call Z4: wide tiles in phase 1 (48-CTA GEMMs) so that the sweep can take more SMs without pushing the 96-CTA GEMMs into two waves
mkdir -p gpurun_out
for v in "56,48:96:256,256" "64,48:96:256,256" "72,48:96:256,256" "64,48:72:256,256" "72,48:72:256,256" "80,48:64:256,256" "64,56:72:256,256" "64,48:84:129,256"; do
  IFS=: read n c w <<< "$v"
  echo "## DVT_FIT_SWEEP_CTAS=$n DVT_FIT_WGRAD_SMS=$c DVT_FIT_X3_WIDE_MIN_N=$w"
  DVT_FIT_SWEEP_CTAS=$n DVT_FIT_WGRAD_SMS=$c DVT_FIT_X3_WIDE_MIN_N=$w timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done | tee gpurun_out/r2z4_sweep_wide.txt
