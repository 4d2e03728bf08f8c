#!/bin/bash
# gpurun call S: new defaults (sweep 48, wgrad SM cap 96, wide tiles in phase 2): tests, variants, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_fit_gpu.py tests/test_stage1_gpu.py -x -q 2>&1 | tail -3
for v in "0,256:96" "0,0:96" "256,256:96" "0,256:80" "0,256:112" "0,512:96" "0,256:64"; do
  w=${v%%:*}; c=${v##*:}
  echo "## DVT_FIT_X3_WIDE_MIN_N=$w DVT_FIT_WGRAD_SMS=$c"
  DVT_FIT_X3_WIDE_MIN_N=$w DVT_FIT_WGRAD_SMS=$c timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done > gpurun_out/r2s_variants.txt 2>&1
cat gpurun_out/r2s_variants.txt
timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"].get("from_image", {}).get("value"))
except Exception as ex:
    print("bench parse failed", ex); print(open("gpurun_out/r2s_bench.err").read()[-2000:])
PY
