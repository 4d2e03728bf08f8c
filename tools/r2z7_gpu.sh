#!/bin/bash
# gpurun call Z7: staggered start of the second CTA per SM; old library on the same box
mkdir -p gpurun_out
{ echo "== old lib (round-2a attention kernel)"; DVT_LIB_PATH=$PWD/denoising-vit_b200/_ab/libdvt_b200_r2a.so timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"
for sg in 0 800 1200 1700 2200 3000; do echo "== DVT_ATTN_STAGGER=$sg"; DVT_ATTN_STAGGER=$sg timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"; done
echo "== non persistent"; DVT_ATTN_PERSISTENT=0 timeout 300 python tools/microbench.py --batch 32 2>&1 | grep "attention"; } | tee gpurun_out/r2z7_attention.txt
