#!/usr/bin/env bash
# 2-GPU validation: bench N=2, stage-1 CLI under torchrun with the NCCL collate, stage-2 DDP bench
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export DVT_ALLOW_RANDOM_INIT=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_2gpu.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])"
# stage-1 CLI, 2 ranks, 5 synthetic images, small ViT-S config
python - <<'PY'
import os, numpy as np
from PIL import Image
os.makedirs('/tmp/dvt_cli/data/set', exist_ok=True)
rs = np.random.RandomState(0)
with open('/tmp/dvt_cli/list.txt', 'w') as f:
    for i in range(5):
        Image.fromarray(rs.randint(0, 255, (90, 120, 3), dtype=np.uint8)).save(f'/tmp/dvt_cli/data/set/{i}.jpg')
        f.write(f'set/{i}.jpg\n')
PY
timeout 600 $TR main_img_denoising.py --model vit_small_patch14_dinov2.lvd142m --input_size 70 84 --stride_size 14 --img_path /tmp/dvt_cli/list.txt --data_root /tmp/dvt_cli/data/ --save_root /tmp/dvt_cli/feats --num_views 6 --num_iters 40 --warmup_iters 4 --n_levels 6 --extract_bsz 4 --pixel_bsz 64 --output_dir /tmp/dvt_cli/work --collate_out /tmp/dvt_cli/maps.pt > gpurun_out/r2k_cli_stage1.log 2>&1; echo "cli stage1 rc=$?"
grep -E "Collated|Total time|Saving" gpurun_out/r2k_cli_stage1.log | tail -4
python - <<'PY'
import torch, glob
p = torch.load('/tmp/dvt_cli/maps.pt')
print('collated', p['raw_feats'].shape, p['denoised_feats'].shape, len(p['files']), 'npy files', len(glob.glob('/tmp/dvt_cli/feats/**/*.npy', recursive=True)))
import numpy as np
for i, f in enumerate(p['files']):
    den = np.load(f.replace('/tmp/dvt_cli/data/', '/tmp/dvt_cli/feats/denoised_features/vit_small_patch14_dinov2.lvd142m/').replace('.jpg', '.npy'))
    assert np.array_equal(den[0], p['denoised_feats'][i].numpy()), i
print('collated stack == .npy store, image order preserved')
PY
# stage-2 CLI from the collated tensors, 2 ranks (one all-reduce per step)
timeout 600 $TR main_denoiser.py --model vit_small_patch14_dinov2.lvd142m --input_size 70 84 --stride_size 14 --collated /tmp/dvt_cli/maps.pt --batch_size 2 --num_iterations 30 --blr 0.02 --output_root /tmp/dvt_cli/work2 --run_name t --save_freq 100 --log_freq 10 > gpurun_out/r2k_cli_stage2.log 2>&1; echo "cli stage2 rc=$?"
grep -E "^Train|Saved" gpurun_out/r2k_cli_stage2.log | tail -5
timeout 600 $TR tools/bench_stage2.py --no-library > gpurun_out/r2k_stage2_2gpu.json 2> gpurun_out/r2k_stage2_2gpu.err; cat gpurun_out/r2k_stage2_2gpu.json
