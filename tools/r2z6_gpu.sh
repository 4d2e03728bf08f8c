#!/bin/bash
# gpurun call Z6: persistent attention CTAs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_gpu.py tests/test_denoiser_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -2
for p in 0 1; do echo "== DVT_ATTN_PERSISTENT=$p"; DVT_ATTN_PERSISTENT=$p timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"; DVT_ATTN_PERSISTENT=$p timeout 300 python tools/microbench.py --batch 16 2>&1 | grep "attention"; done | tee gpurun_out/r2z6_attention.txt
DVT_ATTN_KSTAGES=2 timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention" | tee -a gpurun_out/r2z6_attention.txt
python tools/attention_timeline.py 2>&1 | grep -v "^+" | tail -12 | tee -a gpurun_out/r2z6_attention.txt
