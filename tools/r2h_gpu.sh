#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_stage1_gpu.py -q > gpurun_out/r2h_fit.log 2>&1; tail -4 gpurun_out/r2h_fit.log
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2h_fb_exact.txt 2>&1; tail -1 gpurun_out/r2h_fb_exact.txt
DVT_FIT_EXACT_GRID=0 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2h_fb_plain.txt 2>&1; tail -1 gpurun_out/r2h_fb_plain.txt
DVT_FIT_WGRAD_TF32=1 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2h_fb_wtf32.txt 2>&1; tail -1 gpurun_out/r2h_fb_wtf32.txt
DVT_FIT_WGRAD_TF32=1 DVT_FIT_RES_TF32=1 timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1' > gpurun_out/r2h_fb_wtf32_res.txt 2>&1; tail -1 gpurun_out/r2h_fb_wtf32_res.txt
timeout 600 python tools/diag_early.py 12 2>&1 | grep -v Warning | cut -c1-330 | head -3 > gpurun_out/r2h_early.txt
DVT_FIT_WGRAD_TF32=1 timeout 600 python tools/diag_early.py 12 2>&1 | grep -v Warning | cut -c1-330 | head -1 >> gpurun_out/r2h_early.txt
cat gpurun_out/r2h_early.txt
python - <<'PY' > gpurun_out/r2h_diag.txt 2>&1
import os, sys
sys.argv=['x']
sys.path.insert(0,'tools')
import diag_headline as D
D.run("default", {})
os.environ["DVT_FIT_WGRAD_TF32"]="1"
D.run("wgrad TF32", {"DVT_FIT_WGRAD_TF32":"1"})
D.run("wgrad TF32 + residual TF32", {"DVT_FIT_WGRAD_TF32":"1","DVT_FIT_RES_TF32":"1"})
os.environ.pop("DVT_FIT_WGRAD_TF32"); os.environ.pop("DVT_FIT_RES_TF32", None)
D.run("SIMT fp32 sequential", {"DVT_FIT_PIPELINE":"0"}, impl=1, graph_steps=0)
PY
grep -v Warning gpurun_out/r2h_diag.txt | tail -5
