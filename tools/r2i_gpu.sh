#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
for rep in 1 2; do
DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/old lib, plain cell : /'
DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, plain cell : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, exact grid : /'
done > gpurun_out/r2i_ab.txt
cat gpurun_out/r2i_ab.txt
timeout 300 python tools/microbench.py --batch 32 --only attention > gpurun_out/r2i_attn.txt 2>&1; tail -1 gpurun_out/r2i_attn.txt
timeout 300 python tools/microbench.py --batch 16 --only attention >> gpurun_out/r2i_attn.txt 2>&1; tail -1 gpurun_out/r2i_attn.txt
timeout 600 python -m pytest tests/test_vit_gpu.py tests/test_fit_gpu.py -q -k "attention or golden or headline" > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
timeout 600 python tools/diag_early.py 12 2>&1 | grep -v Warning | cut -c1-330 | sed -n '1p;4p' > gpurun_out/r2i_early.txt; cat gpurun_out/r2i_early.txt
