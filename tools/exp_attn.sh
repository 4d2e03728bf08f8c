set -x
cd $GRAFT_REPO_ROOT
DVT_ATTN_MODE=3 timeout 900 python -m pytest tests/test_vit_gpu.py -x -q 2>&1 | tail -3
for m in 1 3 2; do DVT_ATTN_MODE=$m timeout 300 python tools/microbench.py --only attention --batch 32 2>&1 | tail -1; done
