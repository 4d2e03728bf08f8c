set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -4
timeout 900 python tools/fit_breakdown.py --iters 600 --graphs-only --configs "1:40,-1:20:0;0:0:20:0;1:40,-1:20:1;0:0:20:1;1:40,40:20:1;1:40,32:20:1;1:48,48:50:1" 2>&1 | tail -16
