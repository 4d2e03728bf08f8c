#!/bin/bash
# gpurun call P: does the graph keep priorities / programmatic edges; late W2 weight gradient; wgrad SM caps
mkdir -p gpurun_out
export DVT_FIT_SWEEP_CTAS=48,48
DVT_FIT_DEBUG_GRAPH=1 timeout 600 python tools/fit_breakdown.py --iters 200 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -4
for late in 0 1; do for sms in 148 96 72 48; do
  echo "## DVT_FIT_WGRAD_LATE=$late DVT_FIT_WGRAD_SMS=$sms"
  DVT_FIT_WGRAD_LATE=$late DVT_FIT_WGRAD_SMS=$sms timeout 600 python tools/fit_breakdown.py --iters 600 --graphs-only --graph-steps 20 2>&1 | grep -v "^+" | tail -1
done; done > gpurun_out/r2p_late.txt 2>&1
cat gpurun_out/r2p_late.txt
DVT_FIT_WGRAD_LATE=1 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2p_fit_timeline_late.csv 2>&1 | tail -1
DVT_FIT_WGRAD_LATE=1 DVT_FIT_WGRAD_SMS=72 timeout 600 python tools/fit_timeline.py --out gpurun_out/r2p_fit_timeline_late72.csv 2>&1 | tail -1
DVT_FIT_WGRAD_LATE=1 timeout 900 python -m pytest tests/test_fit_gpu.py -x -q 2>&1 | tail -3
