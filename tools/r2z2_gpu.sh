#!/bin/bash
# gpurun call Z2: the 2-MMA k-step for MN-major B as well (data / weight gradient GEMMs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -3
python tools/gemm_timeline.py 2>&1 | grep event | tee gpurun_out/r2z2_gemm_timeline.txt
for w in "0,256" "0,0" "256,256"; do
echo "## DVT_FIT_X3_WIDE_MIN_N=$w"
DVT_FIT_X3_WIDE_MIN_N=$w timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done | tee gpurun_out/r2z2_fit.txt
