set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tools/fit_breakdown.py --iters 800 --graphs-only --configs "1:40,40:20:1;1:40,40:50:1;1:40,40:100:1;1:32,32:50:1;1:48,40:50:1;1:40,40:20:1" 2>&1 | grep -v "^+" | tail -14
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-kernel-rooflines"
for b in 16 8 24 32; do echo "== extract_bsz $b"; timeout 600 $B --extract-bsz $b | tail -1; done
