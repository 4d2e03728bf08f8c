#!/bin/bash
# gpurun call Z12 (evidence only, closing tree): ncu --set full of the roofline kernels (tools/ncu_targets.py), BASELINE
# configs 2 / 5 + long sequence (tools/bench_extract.py), config 4 (tools/bench_stage2.py)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
R=r2z12
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_targets_$R -f python tools/ncu_targets.py > gpurun_out/${R}_ncu_targets_log.txt 2>&1; tail -1 gpurun_out/${R}_ncu_targets_log.txt
python tools/ncu_summary.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/${R}_targets_ncu_full.txt "tools/ncu_targets.py: 4 ViT-B block GEMMs on CTA pairs (B=16, packed-pair epilogues), dense Adam sweep (full grid, 48 CTAs), fit GEMM F=h1.W2^T (3xTF32), attention fwd(+lse, mode 4) / bwd" | tail -1
python tools/ncu_traffic.py gpurun_out/prof_targets_$R.ncu-rep gpurun_out/traffic.json profiles/${R}_targets_ncu_full.txt | tail -3
rm -f gpurun_out/prof_targets_$R.ncu-rep
echo "elapsed ${SECONDS}s"
timeout 100 python tools/bench_extract.py --reps 3 > gpurun_out/${R}_extract.jsonl 2>/dev/null; cut -c1-220 gpurun_out/${R}_extract.jsonl
echo "elapsed ${SECONDS}s"
if [ $SECONDS -lt 230 ]; then timeout 50 python tools/bench_stage2.py > gpurun_out/${R}_stage2.json 2>/dev/null; cut -c1-300 gpurun_out/${R}_stage2.json; fi
