"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the stage-1 per-image neural-field fit:
  * `denoise_an_image`               main_img_denoising.py:28-149  (loop :67-89, final query :121-130)
  * `SingleImageDenoiser.forward`    dvt/models/offline_denoiser.py:62-171
  * `NeuralFeatureField`             dvt/models/neural_feature_field.py:14-49 (hash grid: oracle/hashgrid.py)
  * `adjust_learning_rate`           dvt/utils/misc.py:306-322
  * `make_patch_coordinates`         main_img_denoising.py:21-25

Reference quirks that are part of the contract and are reproduced here:
  - Adam(lr, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-5) over G, residual MLP, hash table, field MLP, with
    the DENSE hash-table gradient tcnn returns (untouched entries still decay through weight decay);
  - the loss is multiplied by `loss_scale` (GradScaler(2**10) on CUDA, disabled on CPU) and never unscaled;
  - the LR of step `s` is set before the step (step 0 has lr 0 when warm-up > 0);
  - for `step > int(freeze_after * num_iters)`: G stops receiving gradients (torch Adam then SKIPS it: no weight
    decay, no moment update) and the residual MLP starts training, its Adam step counter starting at 1 there.

PINNING: `SingleImageDenoiser` + the Adam loop are checked against the reference's own class, imported from
/root/reference by tests/golden/make_fit_golden.py, which also writes tests/golden/fit_*.npz.  The hash-grid
part inherits "parity unpinned" from oracle/hashgrid.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import hashgrid as HG


def make_patch_coordinates(height: int, width: int, start: float = -1, end: float = 1) -> torch.Tensor:
    """main_img_denoising.py:21-25: (x, y) stacked on the last dim."""
    py, px = torch.linspace(start, end, height), torch.linspace(start, end, width)
    py, px = torch.meshgrid(py, px, indexing="ij")
    return torch.stack([px, py], dim=-1)


def lr_at(step: int, lr: float, min_lr: float, warmup_iters: int, num_iters: int) -> float:
    """dvt/utils/misc.py:306-322."""
    if step < warmup_iters:
        return lr * step / warmup_iters
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (step - warmup_iters) / (num_iters - warmup_iters)))


def init_params(C: int, h: int, w: int, meta: HG.GridMeta, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic initial parameters with the reference's shapes and init scales:
    G = randn * 0.02 (offline_denoiser.py:33-36), nn.Linear default init for the MLPs, hash table U(-1e-4, 1e-4)
    (tcnn default; its own pcg32 stream is not reproducible without tcnn, so the table is an explicit input)."""
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        wt = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound      # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), .)
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return wt, b

    p = {"G": torch.randn(1, C, h, w, generator=g) * 0.02}
    p["res.0.weight"], p["res.0.bias"] = linear(C // 4, C)
    p["res.2.weight"], p["res.2.bias"] = linear(C // 4, C // 4)
    p["res.4.weight"], p["res.4.bias"] = linear(C, C // 4)
    p["table"] = (torch.rand(meta.n_params, generator=g) * 2 - 1) * 1e-4
    p["mlp.0.weight"], p["mlp.0.bias"] = linear(C // 2, meta.n_output_dims)
    p["mlp.2.weight"], p["mlp.2.bias"] = linear(C, C // 2)
    return p


PARAM_ORDER = ["G", "res.0.weight", "res.0.bias", "res.2.weight", "res.2.bias", "res.4.weight", "res.4.bias", "table",
               "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"]  # chain(denoiser, neural_field), :49


def field_forward(p: Dict[str, torch.Tensor], coords: torch.Tensor, meta: HG.GridMeta) -> torch.Tensor:
    """NeuralFeatureField.forward (neural_feature_field.py:46-49)."""
    assert coords.max() <= 1 and coords.min() >= 0, "coordinates should be in [0, 1]"
    enc = HG.encode(p["table"], coords.reshape(-1, 2), meta)
    hid = F.relu(F.linear(enc, p["mlp.0.weight"], p["mlp.0.bias"]))
    return F.linear(hid, p["mlp.2.weight"], p["mlp.2.bias"])


def residual_forward(p: Dict[str, torch.Tensor], raw: torch.Tensor) -> torch.Tensor:
    x = F.relu(F.linear(raw, p["res.0.weight"], p["res.0.bias"]))
    x = F.relu(F.linear(x, p["res.2.weight"], p["res.2.bias"]))
    return F.linear(x, p["res.4.weight"], p["res.4.bias"])


def denoiser_forward(p, raw, global_coords, meta, g_coords, use_residual: bool) -> Dict[str, torch.Tensor]:
    """Training branch of SingleImageDenoiser.forward (offline_denoiser.py:92-140), 2-D inputs."""
    shared = F.grid_sample(p["G"], g_coords[None, None, ...], mode="bilinear", align_corners=True)
    shared = shared.squeeze().permute(1, 0)
    denoised = field_forward(p, global_coords, meta)
    if use_residual:
        pred_res = residual_forward(p, raw)
        pred = denoised + shared + pred_res.detach()
    else:
        pred = shared + denoised
    l2 = F.mse_loss(pred, raw)
    cos = 1 - F.cosine_similarity(pred, raw, dim=-1).mean()
    loss = l2 + cos
    out = {"patch_l2_loss": l2, "cosine_similarity_loss": cos}
    if use_residual:
        gt_res = (raw - denoised - shared).detach()
        rl = 0.1 * F.mse_loss(pred_res, gt_res)
        rs = 0.02 * pred_res.abs().mean()
        loss = loss + rl + rs
        out["residual_loss"], out["residual_sparsity_loss"] = rl, rs
    out["loss"] = loss
    return out


def query(p, raw_hw: torch.Tensor, coords_hw: torch.Tensor, meta, use_residual: bool) -> Dict[str, torch.Tensor]:
    """4-D ("visualization") branch used for the final outputs (offline_denoiser.py:84-91,142-169;
    main_img_denoising.py:121-130): G is used directly as [h*w, C]."""
    shape = raw_hw.shape
    C = shape[-1]
    raw = raw_hw.reshape(-1, C)
    shared = p["G"].permute(0, 2, 3, 1).reshape(-1, C)
    denoised = field_forward(p, coords_hw.reshape(-1, 2), meta)
    out = {"denoised_feats": denoised.reshape(*shape[:-1], -1), "shared_patterns": shared.reshape(*shape[:-1], -1)}
    if use_residual:
        pr = residual_forward(p, raw)
        out["pred_residual"] = pr.reshape(*shape[:-1], -1)
        out["denoised_features"] = (raw - shared - pr).reshape(*shape[:-1], -1)
    else:
        out["denoised_features"] = (raw - shared).reshape(*shape[:-1], -1)
    return out


def fit(bank_feats: torch.Tensor, bank_coords: torch.Tensor, h: int, w: int, meta: HG.GridMeta,
        init: Dict[str, torch.Tensor], idx_stream: np.ndarray, *, lr=0.01, min_lr=0.001, weight_decay=1e-5,
        warmup_iters=200, freeze_after=0.5, loss_scale=1.0, log_every: int = 10) -> Dict[str, object]:
    """bank_feats [V, h, w, C]; bank_coords [V, h, w, 2] (global coords in [0,1]); idx_stream int64 [T, bsz] rows into
    the flattened bank (the np.random.randint stream of main_img_denoising.py:73).  Returns final params + logs."""
    V = bank_feats.shape[0]
    C = bank_feats.shape[-1]
    num_iters = idx_stream.shape[0]
    p = {k: init[k].clone().float().requires_grad_(True) for k in PARAM_ORDER}
    opt = torch.optim.Adam([p[k] for k in PARAM_ORDER], lr=lr, eps=1e-15, weight_decay=weight_decay, betas=(0.9, 0.99))
    g_coords_all = make_patch_coordinates(h, w).unsqueeze(0).repeat(V, 1, 1, 1).reshape(-1, 2)   # :58-62
    feats = bank_feats.reshape(-1, C).float()
    coords = bank_coords.reshape(-1, 2).float()
    use_residual = False
    logs = []
    for step in range(num_iters):
        if step > int(freeze_after * num_iters):                                               # :70-72
            p["G"].requires_grad = False
            use_residual = True
        idx = torch.from_numpy(idx_stream[step].astype(np.int64))
        raw, gc, pc = feats[idx], g_coords_all[idx], coords[idx]
        for grp in opt.param_groups:
            grp["lr"] = lr_at(step, lr, min_lr, warmup_iters, num_iters)
        out = denoiser_forward(p, raw, pc, meta, gc, use_residual)
        opt.zero_grad()
        (out["loss"] * loss_scale).backward()
        opt.step()
        if step % log_every == 0 or step == num_iters - 1:
            logs.append([step] + [float(out[k].detach()) if k in out else 0.0 for k in
                                  ("loss", "patch_l2_loss", "cosine_similarity_loss", "residual_loss",
                                   "residual_sparsity_loss")])
    final = {k: v.detach() for k, v in p.items()}
    with torch.no_grad():
        q = query(final, bank_feats[-1:].float(), bank_coords[-1:].float(), meta, use_residual)
    return {"params": final, "logs": np.array(logs, dtype=np.float64), "denoised_feats": q["denoised_feats"],
            "denoised_features": q["denoised_features"], "use_residual": use_residual}


def synthetic_bank(V: int, h: int, w: int, C: int, seed: int = 0):
    """A bank with the structure DVT assumes: raw = smooth f(global xy) + view-independent artifact G*[r,c] + noise.
    Views are random crops (scale in [0.1, 0.5]) of the unit square, the last view is the full image
    (coords linspace(0,1), main_img_denoising.py:337)."""
    g = torch.Generator().manual_seed(seed)
    freq = torch.randn(2, C, generator=g) * 3.0
    phase = torch.rand(C, generator=g) * 6.28
    amp = 0.5 + torch.rand(C, generator=g)
    g_true = torch.randn(h, w, C, generator=g) * 0.3
    coords = torch.zeros(V, h, w, 2)
    for v in range(V - 1):
        s = (0.1 + 0.4 * torch.rand(1, generator=g).item()) ** 0.5
        x0 = torch.rand(1, generator=g).item() * (1 - s)
        y0 = torch.rand(1, generator=g).item() * (1 - s)
        ys, xs = torch.linspace(y0, y0 + s, h), torch.linspace(x0, x0 + s, w)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        coords[v] = torch.stack([gx, gy], -1)
    coords[-1] = make_patch_coordinates(h, w, 0, 1)
    coords = coords.clamp(0, 1)
    smooth = torch.sin(coords @ freq + phase) * amp
    feats = smooth + g_true.unsqueeze(0) + 0.05 * torch.randn(V, h, w, C, generator=g)
    return feats.float(), coords.float()
