#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_stage1_gpu.py tests/test_gemm_gpu.py tests/test_train_gpu.py -q -s > gpurun_out/r2m_fit.log 2>&1; grep -E "headline golden|passed|failed|Error" gpurun_out/r2m_fit.log | tail -5
{
DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/old lib, plain cell           : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, lean fit epilogues    : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:32,32:20:1;1:48,48:20:1;1:40,40:10:1;1:40,40:40:1' 2>&1 | grep -E "sweep_ctas|phase1" | paste - - | sed 's/^/new lib, other settings: /'
} > gpurun_out/r2m_ab.txt
cat gpurun_out/r2m_ab.txt
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['from_image']['value'], [(o['kernel'][:14], round(o['frac'],3), o.get('ms_per_image')) for o in d['roofline_other']])
PY
