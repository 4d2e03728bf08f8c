#!/bin/bash
# gpurun call V: max-shared carve-out for the fit's small kernels (GEMM CTAs can join their SMs), with both sweep geometries
mkdir -p gpurun_out
for v in "0:1024:48,48" "1:1024:48,48" "1:1024:40,40" "1:1024:56,56" "1:512:148,148" "1:512:96,96" "1:512:148,96" "1:768:64,64" "1:1024:64,64"; do
  IFS=: read c t n <<< "$v"
  echo "## DVT_FIT_CARVEOUT=$c DVT_FIT_SWEEP_THREADS=$t DVT_FIT_SWEEP_CTAS=$n"
  DVT_FIT_CARVEOUT=$c DVT_FIT_SWEEP_THREADS=$t DVT_FIT_SWEEP_CTAS=$n timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done > gpurun_out/r2v_carveout.txt 2>&1
cat gpurun_out/r2v_carveout.txt
DVT_FIT_SWEEP_THREADS=512 DVT_FIT_SWEEP_CTAS=148,148 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2v_fit_timeline_512x148.csv 2>&1 | tail -1
timeout 600 python tools/fit_timeline.py --out gpurun_out/r2v_fit_timeline_default.csv 2>&1 | tail -1
