#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_fit_gpu.py::test_fit_headline_2000_steps_matches_reference_golden > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python tools/diag_headline.py simt > gpurun_out/r2b_diag.txt 2>&1; echo "diag rc=$?"
cat gpurun_out/r2b_diag.txt | tail -12
timeout 300 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20:1;1:32,32:20:1;1:24,24:20:1;1:48,48:20:1' > gpurun_out/r2b_fitbreak.txt 2>&1
cat gpurun_out/r2b_fitbreak.txt | tail -9
