#!/bin/bash
# gpurun call U: one 512-thread (half register file) sweep CTA on every SM, co-resident with the chain's CTAs
mkdir -p gpurun_out
for v in "1024:48,48" "512:148,148" "512:96,96" "512:120,120" "768:64,64" "512:148,96" "256:148,148" "384:148,148"; do
  t=${v%%:*}; c=${v##*:}
  echo "## DVT_FIT_SWEEP_THREADS=$t DVT_FIT_SWEEP_CTAS=$c"
  DVT_FIT_SWEEP_THREADS=$t DVT_FIT_SWEEP_CTAS=$c timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done > gpurun_out/r2u_coresident.txt 2>&1
cat gpurun_out/r2u_coresident.txt
DVT_FIT_SWEEP_THREADS=512 DVT_FIT_SWEEP_CTAS=148,148 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2u_fit_timeline_512x148.csv 2>&1 | tail -1
