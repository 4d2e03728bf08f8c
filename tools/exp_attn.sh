set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vit_gpu.py -x -q 2>&1 | tail -4
DVT_ATTN_LAZY=0 timeout 300 python tools/microbench.py --only attention 2>&1 | tail -3
timeout 300 python tools/microbench.py --only attention 2>&1 | tail -3
timeout 300 python tools/parity_probe.py 2>&1 | tail -4
DVT_FIT_RES_TF32=1 timeout 300 python tools/parity_probe.py 2>&1 | tail -4
timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --configs "1:40,40:20:1" 2>&1 | tail -2
DVT_FIT_RES_TF32=1 timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --configs "1:40,40:20:1" 2>&1 | tail -2
