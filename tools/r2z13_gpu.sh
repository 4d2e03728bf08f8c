#!/bin/bash
# gpurun call Z13 (last GPU minutes of the round): GEMM epilogue fast paths with the scratch reads in front of the stores
# (row chains interleaved): the GPU tests that reach those epilogues, then the per-kernel microbenchmark
# (compare with profiles/r2z10_attention_epilogue_ab.txt: qkv 0.117, fc1 0.176 ms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gemm_gpu.py tests/test_vit_gpu.py tests/test_denoiser_gpu.py -q -x > gpurun_out/r2z13_pytest.log 2>&1; tail -2 gpurun_out/r2z13_pytest.log
timeout 30 python tools/microbench.py --batch 32 --iters 20 2>&1 | grep -v "^+" | tee gpurun_out/r2z13_microbench.txt
