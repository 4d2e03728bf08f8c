#!/bin/bash
# gpurun call Q: 128x128 tiles for the 3xTF32 GEMMs
mkdir -p gpurun_out
export DVT_FIT_SWEEP_CTAS=48,48
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -3
for minn in 0 192 129 256 512; do for sms in 148 96; do
  echo "## DVT_GEMM_X3_WIDE_MIN_N=$minn DVT_FIT_WGRAD_SMS=$sms"
  DVT_GEMM_X3_WIDE_MIN_N=$minn DVT_FIT_WGRAD_SMS=$sms timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done; done > gpurun_out/r2q_wide.txt 2>&1
cat gpurun_out/r2q_wide.txt
DVT_FIT_WGRAD_SMS=96 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2q_fit_timeline_wide96.csv 2>&1 | tail -1
DVT_FIT_WGRAD_SMS=148 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2q_fit_timeline_wide148.csv 2>&1 | tail -1
