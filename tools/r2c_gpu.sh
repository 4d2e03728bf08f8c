#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -s -x > gpurun_out/r2c_train.log 2>&1; echo "train pytest rc=$?" >> gpurun_out/r2c_train.log
tail -25 gpurun_out/r2c_train.log
timeout 600 python -m pytest tests/test_fit_gpu.py -q -k "partial_last_warp or sweep_kernels or device_side" > gpurun_out/r2c_fit.log 2>&1; tail -4 gpurun_out/r2c_fit.log
timeout 600 python tools/diag_early.py 12 > gpurun_out/r2c_early.txt 2>&1
timeout 600 python tools/diag_early.py 60 >> gpurun_out/r2c_early.txt 2>&1
cat gpurun_out/r2c_early.txt | grep -v Warning | tail -8
