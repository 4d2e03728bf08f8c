#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_stage1_gpu.py tests/test_gemm_gpu.py -q -s -x > gpurun_out/r2l_fit.log 2>&1; grep -E "headline golden|passed|failed|Error" gpurun_out/r2l_fit.log | tail -5
{
DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/old lib, plain cell        : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, fused grid scatter : /'
DVT_FIT_FUSE_SCATTER=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, grid-bwd kernel    : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:32,32:20:1;1:48,48:20:1' 2>&1 | grep -E "sweep_ctas|phase1" | paste - - | sed 's/^/new lib, fused, other sweeps: /'
} > gpurun_out/r2l_ab.txt
cat gpurun_out/r2l_ab.txt
timeout 600 python tools/diag_early.py 12 2>&1 | grep -v Warning | cut -c1-330 | sed -n '1p;4p' > gpurun_out/r2l_early.txt; cat gpurun_out/r2l_early.txt
timeout 300 python tools/microbench.py --batch 32 > gpurun_out/r2l_micro.txt 2>&1; cat gpurun_out/r2l_micro.txt
