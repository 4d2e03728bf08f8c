#!/usr/bin/env bash
# Round-2 closing run on one B200: full GPU suite, smoke, ncu captures (top kernels --set full + launch list of the bench
# command), bench line, reference arm.  Outputs under gpurun_out/ (copied into profiles/ by hand).
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
R=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_targets_$R -f python tools/ncu_targets.py > gpurun_out/${R}_ncu_targets_log.txt 2>&1; tail -2 gpurun_out/${R}_ncu_targets_log.txt
python tools/ncu_summary.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/${R}_targets_ncu_full.txt "tools/ncu_targets.py: 4 ViT-B block GEMMs on CTA pairs (B=16), dense Adam sweep (full grid, 48 CTAs), fit GEMM F=h1.W2^T (3xTF32), attention fwd(+lse) / bwd" | tail -1
python tools/ncu_traffic.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/traffic.json profiles/${R}_targets_ncu_full.txt | tail -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_$R.csv python bench.py --steps 1 --warmup 1 --views 37 --num-iters 100 --warmup-iters 10 --no-e2e --no-cpu-baseline --no-kernel-rooflines --no-library-bar > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/launches_bench_$R.csv gpurun_out/${R}_bench_launch_shares.txt "python bench.py --steps 1 --warmup 1 --views 37 --num-iters 100 (reduced: 2 images x (38 views + 100 fit steps))" | head -24
timeout 1200 python bench.py --steps 8 --warmup 3 | tail -1 > gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/${R}_bench_reference.json; cut -c1-300 gpurun_out/${R}_bench_reference.json
timeout 900 python tools/bench_extract.py --reps 3 > gpurun_out/${R}_extract.jsonl 2>/dev/null; cut -c1-200 gpurun_out/${R}_extract.jsonl
timeout 600 python tools/bench_stage2.py > gpurun_out/${R}_stage2.json 2>/dev/null; cut -c1-400 gpurun_out/${R}_stage2.json
