#!/bin/bash
# gpurun call Z9: validation of the restored tree (full GPU suite, smoke, default bench line, launch list of the bench command)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2z9_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2z9_pytest.log 2>&1; tail -3 gpurun_out/r2z9_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2z9_smoke.txt
timeout 600 python bench.py --steps 8 --warmup 3 2>gpurun_out/r2z9_bench.err | tail -1 > gpurun_out/r2z9_bench.json; cut -c1-400 gpurun_out/r2z9_bench.json
