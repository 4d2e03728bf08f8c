#!/bin/bash
# gpurun call X: [hi.hi | hi.lo] as one MMA for K-major B (forward GEMMs of the fit), cheaper GELU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py tests/test_vit_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -3
python tools/gemm_timeline.py 2>&1 | grep event | tee gpurun_out/r2x_gemm_timeline.txt
timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1 | tee gpurun_out/r2x_fit.txt
DVT_FIT_X3_WIDE_MIN_N=256,256 timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1 | tee -a gpurun_out/r2x_fit.txt
timeout 600 python tools/microbench.py --batch 32 2>&1 | grep -v "^+" | head -14 | tee gpurun_out/r2x_microbench.txt
