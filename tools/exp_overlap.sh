set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_stage1_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
echo "== C overlap, sequential sweep"; DVT_FIT_PIPELINE=0 timeout 600 $B | tail -1
echo "== D overlap, default fit schedule"; timeout 600 $B | tail -1
echo "== E overlap, 40,48"; DVT_FIT_SWEEP_CTAS=40,48 timeout 600 $B | tail -1
echo "== G overlap, 64,64"; DVT_FIT_SWEEP_CTAS=64,64 timeout 600 $B | tail -1
echo "== F overlap, default, steps 6 with e2e"; timeout 600 python bench.py --steps 6 --warmup 2 --no-cpu-baseline | tail -1
