#!/bin/bash
# gpurun call Z8: same-box attention A/B: round-2a kernel | committed kernel with timeline stamps | without stamps (default build)
mkdir -p gpurun_out
{ for r in 1 2; do
echo "== old lib (round-2a attention kernel)"; DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"
echo "== with timeline stamps"; DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_timeline.so timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"
echo "== default build"; timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"
echo "== default build, two K stages"; DVT_ATTN_KSTAGES=2 timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"
done; } | tee gpurun_out/r2z8_attention.txt
