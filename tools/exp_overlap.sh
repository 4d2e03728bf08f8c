set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
echo "== persistent ViT GEMM CTAs, no overlap"; timeout 600 $B --no-overlap | tail -1
echo "== persistent ViT GEMM CTAs"; timeout 600 $B | tail -1
for t in 1 2 4; do
echo "== tiles per CTA $t"; DVT_GEMM_TILES_PER_CTA=$t timeout 600 $B | tail -1
done
echo "== tiles per CTA 1, no overlap"; DVT_GEMM_TILES_PER_CTA=1 timeout 600 $B --no-overlap | tail -1
