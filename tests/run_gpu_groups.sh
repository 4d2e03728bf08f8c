#!/usr/bin/env bash
# Runs the GPU test groups in separate processes so that a trapped kernel (sticky CUDA error) in one group does
# not poison the others.  Usage: tests/run_gpu_groups.sh <file> <k-expr> [<k-expr> ...]
f="$1"; shift
mkdir -p gpurun_out
for k in "$@"; do
  echo "=== $f -k '$k'"
  timeout 300 python -m pytest "$f" -m gpu -q -p no:cacheprovider -k "$k" 2>&1 | tail -25
done
