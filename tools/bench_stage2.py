#!/usr/bin/env python
"""BASELINE.json configs[3]: stage-2 training of the generalizable denoiser (one transformer block, ViT-B width) on
synthetic stage-1 outputs, `[16, 37, 37, 768]` per rank, AdamW, data-parallel over N GPUs (one all-reduce of the flat
32.6 MB gradient per step) -- steps/s and samples/s, next to the same step written with torch modules (autograd + SDPA +
cuBLAS + torch.optim.AdamW(fused) in fp32 and under bf16 autocast: the library bar of BASELINE.md 4.5).

  python tools/bench_stage2.py [--steps 30 --warmup 5 --batch 16]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_stage2.py ...
One JSON line on rank 0.  CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "denoising-vit_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def time_steps(step, steps, warmup, world):
    import torch.distributed as dist
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-library", action="store_true")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import dvt.models as DVT
    from dvt import _lib, train_ops
    from dvt.optim import FusedAdamW
    B, h, w, C = a.batch, 37, 37, 768
    g = torch.Generator(device=dev).manual_seed(rank)
    feats = torch.randn(B, h, w, C, device=dev, generator=g)
    target = feats + 0.1 * torch.randn(B, h, w, C, device=dev, generator=g)
    model = DVT.Denoiser(h, w, C, vit=None, num_blocks=1).to(dev).train()
    opt = FusedAdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-5)

    def step():
        loss, _, _ = train_ops.denoise_loss(model(feats), target)
        opt.zero_grad()
        loss.backward()
        opt.sync_grads(world)
        opt.step()

    l0 = _lib.lib().dvt_launch_count()
    ms = time_steps(step, a.steps, a.warmup, world)
    launches = (_lib.lib().dvt_launch_count() - l0) / (a.steps + a.warmup)
    # algorithmic FLOPs of one block fwd + bwd on B x 1369 tokens (2 M N K; backward = 2 x forward for the GEMMs, 2.5 x for
    # attention with its recomputation)
    M, N = B * h * w, h * w
    gemm = 2 * M * C * (3 * C + C + 4 * C + 4 * C)
    attn = 4 * B * 12 * N * N * 64
    flops = 3 * gemm + 3.5 * attn
    out = {"metric": "stage-2 training steps/s (Denoiser, 1 block, ViT-B width)", "value": world * 1000.0 / ms / world, "unit": "steps/s",
           "samples_per_s": world * B * 1000.0 / ms, "n_gpus": world, "ms_per_step": ms, "batch_per_gpu": B,
           "gpu_launches_per_step": launches, "tflops_per_gpu": flops / (ms / 1e3) / 1e12, "dtype": "bf16 GEMMs / attention, fp32 master weights and residual stream",
           "data": "synthetic", "scaling": "weak", "allreduce_bytes_per_step": int(opt.numel * 4) if world > 1 else 0}
    if not a.no_library and rank == 0:
        import library_bar
        ref = library_bar._Block(C, 12).to(dev).train()
        pos = torch.nn.Parameter(torch.zeros(1, h * w, C, device=dev))
        ropt = torch.optim.AdamW(list(ref.parameters()) + [pos], lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-5, fused=True)
        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            def rstep():
                with ctx:
                    pred = ref(feats.reshape(B, h * w, C) + pos).reshape(B, h, w, C).float()
                    loss = F.mse_loss(pred, target) + 1 - F.cosine_similarity(pred, target, dim=-1).mean()
                ropt.zero_grad()
                loss.backward()
                ropt.step()
            out[f"library_bar_ms_per_step_{name}"] = time_steps(rstep, max(3, a.steps // 3), 2, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
