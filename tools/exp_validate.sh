set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/microbench.py 2>&1 | tail -14
DVT_VIT_PDL=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline | tail -1 > gpurun_out/bench_r1x_nopdl.json
timeout 900 python bench.py --steps 3 --warmup 3 | tail -1 > gpurun_out/bench_r1x.json; cat gpurun_out/bench_r1x.json | cut -c1-300
