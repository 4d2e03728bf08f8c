set -x
cd $GRAFT_REPO_ROOT
R=${1:-r1y}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_targets_$R -f python tools/ncu_targets.py > gpurun_out/ncu_targets_log.txt 2>&1; tail -2 gpurun_out/ncu_targets_log.txt
python tools/ncu_summary.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/${R}_targets_ncu_full.txt "tools/ncu_targets.py: 4 ViT-B block GEMMs (B=16), dense Adam sweep, fit GEMM F=h1.W2^T (3xTF32)" | tail -1
python tools/ncu_traffic.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/traffic.json profiles/${R}_targets_ncu_full.txt | tail -30
cp gpurun_out/traffic.json profiles/traffic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --views 37 --num-iters 100 --warmup-iters 10 --no-e2e --no-cpu-baseline --no-kernel-rooflines > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/launches_bench_$R.csv gpurun_out/launch_shares_$R.txt | head -30
timeout 900 python bench.py --steps 3 --warmup 3 | tail -1 > gpurun_out/bench_$R.json; cut -c1-250 gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 | tail -1 | cut -c1-400
