"""Host-side helpers of the stage-1 / stage-2 drivers with the reference's names and behaviour
(reference dvt/utils/misc.py: fix_random_seeds :19-23, adjust_learning_rate :306-322, check_if_file_exists :325-337).
Only what the hot-path drivers call is provided."""
from __future__ import annotations

import math
import os
import random

import numpy as np
import torch


def fix_random_seeds(seed: int = 31):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def learning_rate_at(iteration: int, lr: float, min_lr: float, warmup_iters: int, num_iters: int) -> float:
    """Linear warm-up then half-cosine decay to min_lr (the schedule the fit engine precomputes per step)."""
    if iteration < warmup_iters:
        return lr * iteration / warmup_iters
    t = (iteration - warmup_iters) / (num_iters - warmup_iters)
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * t))


def adjust_learning_rate(optimizer, iteration, args):
    """Same signature as the reference: sets param_group['lr'] (honouring 'lr_scale') and returns lr."""
    lr = learning_rate_at(iteration, args.lr, args.min_lr, args.warmup_iters, args.num_iters)
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr


def feature_paths(args, filename: str):
    """(raw_path, denoised_path) of the stage-1 outputs for an image path (reference main_img_denoising.py:131-139)."""
    ext = os.path.splitext(filename)[1]
    raw_dir = f"{args.save_root}/raw_features/{args.model}/"
    den_dir = f"{args.save_root}/denoised_features/{args.model}/"
    return (filename.replace(args.data_root, raw_dir).replace(ext, ".npy"),
            filename.replace(args.data_root, den_dir).replace(ext, ".npy"))


def check_if_file_exists(args, filename: str) -> bool:
    raw_path, den_path = feature_paths(args, filename)
    return os.path.isfile(raw_path) and os.path.isfile(den_path)


def cosine_schedule(it: int, base_value: float, final_value: float, total_iters: int, warmup_iters: int = 0,
                    start_warmup_value: float = 0.0) -> float:
    """Value at iteration `it` of the stage-2 schedule (reference `CosineScheduler`, dvt/utils/misc.py:211-241): a linear
    ramp of `warmup_iters` points from start_warmup_value to base_value INCLUSIVE (np.linspace), then a half cosine over
    the remaining iterations that reaches final_value only after the last one; final_value from total_iters on."""
    if it >= total_iters:
        return final_value
    if it < warmup_iters:
        return float(np.linspace(start_warmup_value, base_value, warmup_iters)[it])
    n = total_iters - warmup_iters
    return float(final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * (it - warmup_iters) / n)))


def apply_optim_scheduler(optimizer, lr: float):
    for group in optimizer.param_groups:
        group["lr"] = lr
