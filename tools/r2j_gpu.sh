#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
{
DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/old lib, plain cell : /'
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, exact grid : /'
DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' 2>&1 | tail -1 | sed 's/^/new lib, plain cell : /'
} > gpurun_out/r2j_ab.txt
cat gpurun_out/r2j_ab.txt
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_stage1_gpu.py -q -s > gpurun_out/r2j_fit.log 2>&1; grep -E "headline golden|passed|failed" gpurun_out/r2j_fit.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['from_image']['value'], 'roofline', d['roofline']['frac'], d['roofline']['full_grid']['frac'], [(o['kernel'][:14], round(o['frac'],3), o.get('ms_per_image')) for o in d['roofline_other']], d['library_bar']['fit_wall_clock_ratio'], d['library_bar']['images_per_s_ratio'])
PY
