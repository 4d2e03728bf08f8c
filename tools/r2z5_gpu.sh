#!/bin/bash
# gpurun call Z5: attention with two K stages, faster im2col
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_gpu.py tests/test_denoiser_gpu.py tests/test_stage1_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -2
for k in 1 2; do echo "== DVT_ATTN_KSTAGES=$k"; DVT_ATTN_KSTAGES=$k timeout 300 python tools/microbench.py --batch 32 2>&1 | grep -v "^+" | head -10; done | tee gpurun_out/r2z5_microbench.txt
DVT_ATTN_KSTAGES=2 DVT_ATTN_MODE=3 timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention" | tee -a gpurun_out/r2z5_microbench.txt
timeout 600 python tools/bench_extract.py --reps 3 2>/dev/null | cut -c1-330 | tee gpurun_out/r2z5_extract.txt
for st in 3 4; do echo "## DVT_GEMM_X3_STAGES=$st"; DVT_GEMM_X3_STAGES=$st timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1; done | tee gpurun_out/r2z5_x3stages.txt
DVT_GEMM_X3_STAGES=4 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -2
