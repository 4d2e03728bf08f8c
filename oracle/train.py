"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the stage-2 TRAINING step:
  * model          `Denoiser.forward`, dvt/models/online_denoiser.py:62-104 -> oracle/denoiser.py::forward (autograd on)
  * objective      main_denoiser.py:214-217: F.mse_loss + (1 - F.cosine_similarity(dim=-1).mean())
  * optimiser      main_denoiser.py:176-180: torch.optim.AdamW(betas=(0.9, 0.999), weight_decay), lr set per step from
                   `CosineScheduler` (dvt/utils/misc.py:211-241) through `apply_optim_scheduler`
Gradients come from torch autograd, the update from torch.optim.AdamW itself.

PINNING: the schedule is pinned to the reference's own `CosineScheduler` (tests/golden/store_and_schedule.json, minted by
tests/golden/make_store_golden.py); the block inherits the pinning of oracle/vit.py (transformers Dinov2 fixtures); timm's
`Block` itself is not installable here -> "parity unpinned" against timm for the block, as for oracle/denoiser.py."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import denoiser as OD


def cosine_schedule(base_value: float, final_value: float, total_iters: int, warmup_iters: int = 0,
                    start_warmup_value: float = 0.0) -> np.ndarray:
    """dvt/utils/misc.py:211-235 (freeze_iters = 0)."""
    warm = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(total_iters - warmup_iters)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    out = np.concatenate((warm, sched))
    assert len(out) == total_iters
    return out


def loss_terms(pred: torch.Tensor, target: torch.Tensor):
    l2 = F.mse_loss(pred, target)
    cos = 1 - F.cosine_similarity(pred, target, dim=-1).mean()
    return l2 + cos, l2, cos


def gradients(sd: Dict[str, torch.Tensor], x: torch.Tensor, target: torch.Tensor, hw: Tuple[int, int], num_blocks: int = 1):
    """(loss terms, {name: grad}) of one forward / backward at parameters `sd` (x, target: [b, h, w, c])."""
    p = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
    xin = x.clone().float().requires_grad_(True)
    pred = OD.forward(p, xin, hw, num_blocks)
    loss, l2, cos = loss_terms(pred, target.float())
    loss.backward()
    grads = {k: v.grad.clone() for k, v in p.items()}
    grads["__input__"] = xin.grad.clone()
    return (float(loss.detach()), float(l2.detach()), float(cos.detach())), grads, pred.detach()


def train(sd: Dict[str, torch.Tensor], batches: List[Tuple[torch.Tensor, torch.Tensor]], hw: Tuple[int, int], *, lr_values,
          weight_decay: float = 1e-5, num_blocks: int = 1):
    """Runs len(batches) AdamW steps; returns (final state dict, per-step (loss, l2, cos))."""
    p = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(p.values()), betas=(0.9, 0.999), weight_decay=weight_decay)
    logs = []
    for step, (x, t) in enumerate(batches):
        for g in opt.param_groups:
            g["lr"] = float(lr_values[step])
        pred = OD.forward(p, x.float(), hw, num_blocks)
        loss, l2, cos = loss_terms(pred, t.float())
        opt.zero_grad()
        loss.backward()
        opt.step()
        logs.append((float(loss.detach()), float(l2.detach()), float(cos.detach())))
    return {k: v.detach() for k, v in p.items()}, np.array(logs)
