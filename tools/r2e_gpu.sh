#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/r2e_gemm.log 2>&1; echo "gemm pytest rc=$?" >> gpurun_out/r2e_gemm.log
tail -12 gpurun_out/r2e_gemm.log
if grep -q "rc=0" gpurun_out/r2e_gemm.log; then
  timeout 300 python tools/microbench.py --batch 32 > gpurun_out/r2e_micro_cg2.txt 2>&1
  DVT_GEMM_CG2=0 timeout 300 python tools/microbench.py --batch 32 > gpurun_out/r2e_micro_1cta.txt 2>&1
  paste gpurun_out/r2e_micro_cg2.txt gpurun_out/r2e_micro_1cta.txt | cut -c1-140
else
  export DVT_GEMM_CG2=0
fi
DIAG_ELEMENTWISE=1 timeout 600 python tools/diag_early.py 3 2>&1 | grep -v Warning | cut -c1-700 > gpurun_out/r2e_early.txt
timeout 600 python tools/diag_early.py 12 2>&1 | grep -v Warning | cut -c1-400 >> gpurun_out/r2e_early.txt
cat gpurun_out/r2e_early.txt
timeout 900 python tools/diag_headline.py > gpurun_out/r2e_diag.txt 2>&1; tail -8 gpurun_out/r2e_diag.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_fit_gpu.py::test_fit_headline_2000_steps_matches_reference_golden > gpurun_out/r2e_pytest.log 2>&1; tail -8 gpurun_out/r2e_pytest.log
