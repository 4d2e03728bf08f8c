#!/usr/bin/env python
"""Prints the parity figures of the fit against the committed reference goldens (tests/golden/fit_*.npz) for the
engine configuration selected through the environment (DVT_FIT_* knobs): min cosine of `denoised_feats` and of
`denoised_features = raw - G - R`, max |G - G_ref|.  Used to put numbers on schedule / precision experiments."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_fit_gpu as T  # noqa: E402
from dvt.fit import FitEngine  # noqa: E402


def min_cos(a, b):
    return F.cosine_similarity(a.float().reshape(-1, a.shape[-1]), b.float().reshape(-1, b.shape[-1]), dim=-1).min().item()


def main():
    knobs = {k: v for k, v in os.environ.items() if k.startswith("DVT_")}
    print("knobs:", knobs)
    for name in ("small_L6", "small_L6_ls1", "hashed_L16"):
        cfg, z = T._golden(name)
        feats, coords, init, idx, den, field, _ = T._setup(cfg)
        eng = FitEngine(cfg["C"], cfg["h"], cfg["w"], cfg["bsz"], field.meta)
        eng.fit(den, field, feats.reshape(-1, cfg["C"]).cuda().contiguous(), coords.reshape(-1, 2).cuda().contiguous(), idx,
                graph_steps=7, lr=cfg["lr"], min_lr=cfg["min_lr"], warmup_iters=cfg["warmup_iters"],
                freeze_after=cfg["freeze_after"], weight_decay=cfg["weight_decay"], loss_scale=cfg["loss_scale"])
        den_f = eng.query(coords[-1:].cuda()).cpu()
        resid = eng.residual(feats[-1:].cuda()).cpu()
        G = eng.get_param("G", den.shared_artifacts).permute(0, 2, 3, 1).cpu()
        clean = feats[-1:] - G - resid
        c1 = min_cos(den_f, torch.from_numpy(z["denoised_feats"]))
        c2 = min_cos(clean, torch.from_numpy(z["denoised_features"]))
        gd = (G.permute(0, 3, 1, 2) - torch.from_numpy(z["G_final"])).abs().max().item()
        losses = eng.losses()
        lrel = max(abs(losses[int(r[0]), 0] - r[1]) / (abs(r[1]) + 1e-9) for r in z["logs"])
        print(f"{name:14s} 1-cos(denoised_feats) {1 - c1:.3e}   1-cos(denoised_features) {1 - c2:.3e}   max|G-Gref| {gd:.3e}"
              f"   max rel loss err {lrel:.3e}")


if __name__ == "__main__":
    main()
