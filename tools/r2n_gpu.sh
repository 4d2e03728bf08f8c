#!/bin/bash
# gpurun call N: fit kernel timeline (CUPTI) + sweep CTA counts with the lean epilogues
mkdir -p gpurun_out
timeout 600 python tools/fit_timeline.py --out gpurun_out/r2n_fit_timeline.csv > gpurun_out/r2n_timeline.log 2>&1
tail -3 gpurun_out/r2n_timeline.log
timeout 900 python tools/fit_breakdown.py --iters 600 --graphs-only --configs '1:40,40:20;1:48,48:20;1:56,56:20;1:64,64:20;1:56,48:20;1:48,40:20;1:64,56:20' 2>&1 | grep -v "^+" | paste - - > gpurun_out/r2n_sweeps.txt
cat gpurun_out/r2n_sweeps.txt
