#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2g_fitbreak.txt 2>&1; tail -1 gpurun_out/r2g_fitbreak.txt
DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2g_fitbreak_plain.txt 2>&1; tail -1 gpurun_out/r2g_fitbreak_plain.txt
timeout 300 python -m pytest tests/test_stage1_gpu.py -q -x > gpurun_out/r2g_stage1.log 2>&1; tail -3 gpurun_out/r2g_stage1.log
for ne in 2 1; do
  DVT_FIT_ENGINES=$ne timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar --no-kernel-rooflines > gpurun_out/r2g_bench_ne$ne.json 2> gpurun_out/r2g_bench_ne$ne.err; echo "bench ne=$ne rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench_ne$ne.json').read().strip().splitlines()[-1])
print('engines $ne', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['from_image']['value'], [(o['kernel'][:12], o.get('ms_per_image')) for o in d['roofline_other'][1:]])
PY
done
DVT_FIT_ENGINES=2 DVT_FIT_SWEEP_CTAS=24,24 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar --no-kernel-rooflines --no-e2e > gpurun_out/r2g_bench_ne2_s24.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_ne2_s24.json').read().strip().splitlines()[-1]); print('engines 2 sweep 24', d['value'], d['ms_per_step'])"
DVT_FIT_ENGINES=3 DVT_FIT_SWEEP_CTAS=32,32 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar --no-kernel-rooflines --no-e2e > gpurun_out/r2g_bench_ne3_s32.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_ne3_s32.json').read().strip().splitlines()[-1]); print('engines 3 sweep 32', d['value'], d['ms_per_step'])"
