#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r2d_train.log 2>&1; echo "train pytest rc=$?" >> gpurun_out/r2d_train.log
tail -6 gpurun_out/r2d_train.log
DIAG_ELEMENTWISE=1 timeout 600 python tools/diag_early.py 3 > gpurun_out/r2d_early.txt 2>&1
DIAG_ELEMENTWISE=1 timeout 600 python tools/diag_early.py 12 >> gpurun_out/r2d_early.txt 2>&1
grep -v Warning gpurun_out/r2d_early.txt | tail -20
