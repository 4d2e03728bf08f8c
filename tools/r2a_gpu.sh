#!/usr/bin/env bash
# round 2, GPU call A: full GPU test suite, stage timings, bench line, configs 2/5 extraction bench
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2a_fitbreak.txt 2>&1
tail -3 gpurun_out/r2a_fitbreak.txt
timeout 600 python bench.py --steps 4 --warmup 2 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
timeout 900 python tools/bench_extract.py --reps 3 > gpurun_out/r2a_extract.jsonl 2> gpurun_out/r2a_extract.err; echo "extract rc=$?"
cat gpurun_out/r2a_extract.jsonl | cut -c1-400
