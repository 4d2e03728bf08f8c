#!/bin/bash
# gpurun call Z3: sweep CTAs again (phase 1 is sweep-bound now), wgrad cap with it
mkdir -p gpurun_out
for v in "48,48:96" "52,48:96" "56,48:96" "60,48:96" "64,48:96" "56,48:88" "56,52:96" "64,56:88"; do
  n=${v%%:*}; c=${v##*:}
  echo "## DVT_FIT_SWEEP_CTAS=$n DVT_FIT_WGRAD_SMS=$c"
  DVT_FIT_SWEEP_CTAS=$n DVT_FIT_WGRAD_SMS=$c timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done | tee gpurun_out/r2z3_sweep.txt
