set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_fit_gpu.py -x -q 2>&1 | tail -8
timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --configs "0:0:20;1:48,32:20;1:48,32:50;1:40,24:50;1:56,40:50;1:64,48:50;1:32,20:50;1:0:50" 2>&1 | tail -30
