#!/bin/bash
# gpurun call T: separable view_crops, bench twice (from_image variance)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stage1_gpu.py -x -q 2>&1 | tail -3
python tools/bench_views.py 2>&1 | grep view_crops | tee gpurun_out/r2t_views.txt
for i in 1 2; do
timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2t_bench$i.json 2> gpurun_out/r2t_bench$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2t_bench$i.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"].get("from_image", {}).get("value"))
PY
done
