#!/bin/bash
# gpurun call R: is the x3 mainloop MMA-instruction bound (mode 2 = a third of the MMAs)?  streaming hints in the sweep
mkdir -p gpurun_out
for m in 1 2; do echo "== BN=64 DVT_DEBUG_X3_MODE=$m"; DVT_GEMM_X3_WIDE_MIN_N=0 DVT_DEBUG_X3_MODE=$m python tools/gemm_timeline.py 2>&1 | grep event; done
export DVT_FIT_SWEEP_CTAS=48,48 DVT_FIT_WGRAD_SMS=96
for st in 0 1; do for minn in 0 256 192; do
  echo "## DVT_FIT_SWEEP_STREAM=$st DVT_GEMM_X3_WIDE_MIN_N=$minn"
  DVT_FIT_SWEEP_STREAM=$st DVT_GEMM_X3_WIDE_MIN_N=$minn timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done; done > gpurun_out/r2r_stream.txt 2>&1
cat gpurun_out/r2r_stream.txt
DVT_GEMM_X3_WIDE_MIN_N=256 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2r_fit_timeline.csv 2>&1 | tail -1
timeout 600 python -m pytest tests/test_fit_gpu.py -x -q -k "sweep or golden" 2>&1 | tail -2
