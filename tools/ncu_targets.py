#!/usr/bin/env python
"""One launch each of the kernels bench.py reports rooflines for, between cudaProfilerStart / Stop, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o <rep> python tools/ncu_targets.py
Order of the captured launches: GEMM qkv, out-proj(+residual), fc1(+GELU), fc2(+residual) of one ViT-B block at
batch 16 (gemm_tn_cg2_kernel: bf16, CTA pairs), then the dense Adam sweep of the full-size hash table on the full grid and
on 48 persistent CTAs (fit_adam_table_kernel, the geometry of the timed region), the fit's largest 3xTF32 GEMM
(F = h1 W2^T: M 2048, N 768, K 384), flash attention forward (+lse) and backward at batch 16 x 12 heads x 1370 tokens.
tools/ncu_traffic.py turns the report into profiles/traffic.json."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
import dvt.models as DVT  # noqa: E402
from dvt import ops, train_ops  # noqa: E402
from dvt.fit import FitEngine  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    B, N, C = 16, 1370, 768
    M = B * N
    x, xn, hid = rn(M, C), rn(M, C).bfloat16(), rn(M, 4 * C).bfloat16()
    w_qkv, b_qkv = (rn(3 * C, C) / 28).bfloat16(), rn(3 * C)
    w_proj, b_proj = (rn(C, C) / 28).bfloat16(), rn(C)
    w_fc1, b_fc1 = (rn(4 * C, C) / 28).bfloat16(), rn(4 * C)
    w_fc2, b_fc2 = (rn(C, 4 * C) / 55).bfloat16(), rn(C)
    gam = torch.full((C,), 1e-3, device=dev)
    gemms = [lambda: ops.gemm_tn(xn, w_qkv, b_qkv, None, torch.bfloat16),
             lambda: ops.gemm_tn_residual_(x, xn, w_proj, b_proj, gam),
             lambda: ops.gemm_tn(xn, w_fc1, b_fc1, "gelu", torch.bfloat16),
             lambda: ops.gemm_tn_residual_(x, hid, w_fc2, b_fc2, gam)]
    # full-size fit engine on a small synthetic bank
    h = w = 37
    V, bsz, iters = 8, 2048, 6
    field = DVT.NeuralFeatureField(feat_dim=C, n_levels=16)
    den = DVT.SingleImageDenoiser(h, w, C)
    eng = FitEngine(C, h, w, bsz, field.meta)
    bank = rn(V * h * w, C)
    coords = torch.rand(V * h * w, 2, device=dev, generator=g)
    idx = np.random.RandomState(0).randint(0, V * h * w, (iters, bsz))
    eng.fit(den, field, bank, coords, idx, graph_steps=0, lr=0.01, min_lr=0.001, warmup_iters=2, freeze_after=0.5,
            weight_decay=1e-5, loss_scale=1024.0)
    a2, b2 = rn(2048, 384), rn(768, 384)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    qkv = rn(B, N, 3 * C).bfloat16()
    att, lse = train_ops.attention_fwd_lse(qkv, 12)
    datt = rn(B, N, C).bfloat16()
    targets = gemms + [lambda: eng.sweep_once(0), lambda: eng.sweep_once(-48), lambda: ops.gemm_f32x3(a2, b2, 2048, 768, 384),
                       lambda: train_ops.attention_fwd_lse(qkv, 12), lambda: train_ops.attention_bwd(qkv, att, datt, lse, 12)]
    for fn in targets:          # warm-up (module load, attribute opt-in)
        fn()
    torch.cuda.synchronize()
    for fn in targets:
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        fn()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("captured", len(targets), "launches")


if __name__ == "__main__":
    main()
