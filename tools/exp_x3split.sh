set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py tests/test_stage1_gpu.py -x -q 2>&1 | tail -12
timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --configs "-:40,-1:20;0:0:20;-:40,40:20" 2>&1 | tail -8
timeout 300 python tools/gemm_timeline.py 2>&1 | tail -12
