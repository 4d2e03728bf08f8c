set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vit_gpu.py -x -q -k attention 2>&1 | tail -2
for m in 1 2; do DVT_ATTN_MODE=$m timeout 300 python tools/microbench.py --only attention --batch 32 2>&1 | tail -1; done
