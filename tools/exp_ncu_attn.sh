set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/prof_attn_r1x -f python tools/microbench.py --only attention --iters 2 > gpurun_out/ncu_attn_log.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_attn_r1x.ncu-rep gpurun_out/r1x_attention_ncu_full.txt "attention LAZY (B=16 N=1370 H=12 d=64)"
cat gpurun_out/r1x_attention_ncu_full.txt
