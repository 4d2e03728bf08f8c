#!/bin/bash
# gpurun call Z11 (closing run of round 2): attention mode 6 (P stored chunk by chunk) against the new default 4, then the
# full GPU suite, smoke and the bench line of the tree as committed; ncu capture of the attention kernel; launch list of the
# (reduced) bench command if time is left.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
O=gpurun_out/r2z11_ab.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2z11_smi.txt 2>&1
DVT_ATTN_MODE=6 timeout 240 python -m pytest tests/test_vit_gpu.py tests/test_train_gpu.py tests/test_denoiser_gpu.py -q -x > gpurun_out/r2z11_mode6.log 2>&1
echo "mode 6 tests rc=$? : $(tail -1 gpurun_out/r2z11_mode6.log)" | tee -a $O
for m in 4 6 4 6; do
  echo "== DVT_ATTN_MODE=$m" >> $O
  DVT_ATTN_MODE=$m timeout 120 python tools/microbench.py --batch 32 --iters 30 --only attention 2>&1 | grep "^attention" >> $O
done
cat $O
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2z11_pytest.log 2>&1; tail -2 gpurun_out/r2z11_pytest.log | tee -a $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2z11_smoke.txt
timeout 400 python bench.py --steps 8 --warmup 3 2>gpurun_out/r2z11_bench.err | tail -1 > gpurun_out/r2z11_bench.json; cut -c1-330 gpurun_out/r2z11_bench.json
BEST=$(python - <<'PY'
import re
t = open("gpurun_out/r2z11_ab.txt").read()
ok6 = re.search(r"mode 6 tests rc=0", t) is not None
ms = {4: [], 6: []}
for m, v in re.findall(r"== DVT_ATTN_MODE=(\d)\nattention\s+([\d.]+)", t):
    ms[int(m)].append(float(v))
b = 4
if ok6 and ms[4] and ms[6] and max(ms[6]) < 0.98 * min(ms[4]):
    b = 6
print(b)
PY
)
echo "best attention mode: $BEST (elapsed ${SECONDS}s)" | tee -a $O
if [ "$BEST" = 6 ]; then
  export DVT_ATTN_MODE=6
  timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2z11_pytest_mode6.log 2>&1; tail -1 gpurun_out/r2z11_pytest_mode6.log | tee -a $O
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar --no-kernel-rooflines 2>/dev/null | tail -1 > gpurun_out/r2z11_bench_mode6.json; cut -c1-200 gpurun_out/r2z11_bench_mode6.json
fi
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/prof_attn_r2z11 -f python tools/microbench.py --batch 16 --iters 1 --only attention > gpurun_out/r2z11_ncu_attn_log.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_attn_r2z11.ncu-rep gpurun_out/r2z11_attention_ncu_full.txt "attention forward, DVT_ATTN_MODE=$BEST (B=16 N=1370 H=12 d=64), tools/microbench.py --batch 16 --only attention" | tail -1
rm -f gpurun_out/prof_attn_r2z11.ncu-rep
echo "elapsed before the launch list: ${SECONDS}s" | tee -a $O
if [ $SECONDS -lt 400 ]; then
  timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r2z11.csv python bench.py --steps 1 --warmup 1 --views 37 --num-iters 100 --warmup-iters 10 --no-e2e --no-cpu-baseline --no-kernel-rooflines --no-library-bar > /dev/null 2>&1
  python tools/launch_shares.py gpurun_out/launches_bench_r2z11.csv gpurun_out/r2z11_bench_launch_shares.txt "python bench.py --steps 1 --warmup 1 --views 37 --num-iters 100 (reduced: 2 images x (38 views + 100 fit steps))" | head -12
fi
