#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 300 python tools/microbench.py --batch 32 > gpurun_out/r2f_micro.txt 2>&1; cat gpurun_out/r2f_micro.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vit_gpu.py tests/test_denoiser_gpu.py tests/test_train_gpu.py -q > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 600 python -m pytest tests/test_fit_gpu.py -q -k "headline or golden" -s > gpurun_out/r2f_fit.log 2>&1; grep -E "headline golden|passed|failed" gpurun_out/r2f_fit.log
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2f_fitbreak.txt 2>&1; tail -2 gpurun_out/r2f_fitbreak.txt
timeout 900 python bench.py --steps 5 --warmup 2 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], [ (o['kernel'][:30], round(o['frac'],3), o.get('ms_per_image')) for o in d['roofline_other']], d.get('library_bar',{}).get('fit_wall_clock_ratio'))
PY
timeout 900 python tools/bench_extract.py --reps 3 --no-library > gpurun_out/r2f_extract.jsonl 2> gpurun_out/r2f_extract.err; cut -c1-330 gpurun_out/r2f_extract.jsonl
timeout 600 python tools/bench_stage2.py > gpurun_out/r2f_stage2.json 2> gpurun_out/r2f_stage2.err; cat gpurun_out/r2f_stage2.json; tail -3 gpurun_out/r2f_stage2.err
