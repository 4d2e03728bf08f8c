"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's view generation:
  * `RandomResizedCropFlip.forward`     dvt/dataset/transform.py:39-76
  * the sampling of its parameters      torchvision `RandomResizedCrop.get_params` (torch RNG) and the flip decision
                                        `np.random.random() < 0.5` (numpy RNG), transform.py:48,69
The resampling itself is torchvision 0.26 `F.resized_crop(..., BICUBIC, antialias=True)` on tensors (ATen
`_upsample_bicubic2d_aa`), the library the reference calls; it is used here as is.

PINNING: tests/golden/make_views_golden.py imports the reference's own class from /root/reference, runs it with seeded
RNGs and stores views + coordinates (tests/golden/views_small.npz); tests/test_oracle_views.py checks this restatement
against them bit for bit."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torchvision
import torchvision.transforms.functional as TF


def sample_view_params(img_chw: torch.Tensor, num_views: int, scale=(0.1, 0.5), ratio=(3.0 / 4.0, 4.0 / 3.0),
                       horizontal_flip: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """The RNG calls of `num_views` successive `forward` calls (transform.py:48 then :69): boxes int32 [V, 4] =
    (top, left, height, width), flips int32 [V]."""
    boxes, flips = [], []
    for _ in range(num_views):
        boxes.append(torchvision.transforms.RandomResizedCrop.get_params(img_chw, list(scale), list(ratio)))
        flips.append(int(horizontal_flip and np.random.random() < 0.5))
    return np.asarray(boxes, dtype=np.int32).reshape(-1, 4), np.asarray(flips, dtype=np.int32)


def patch_grid(size, patch_size: int, stride: int) -> Tuple[int, int]:
    """transform.py:36-37"""
    return (size[0] - patch_size) // stride + 1, (size[1] - patch_size) // stride + 1


def make_views(img_chw: torch.Tensor, boxes: np.ndarray, flips: np.ndarray, size, patch_size: int = 14, stride: int = 14):
    """transform.py:47-76 for given parameters: views [V, 3, OH, OW] and coords [V, hp, wp, 2] (x, y)."""
    _, H, W = img_chw.shape
    hp, wp = patch_grid(size, patch_size, stride)
    views: List[torch.Tensor] = []
    coords: List[torch.Tensor] = []
    for (i, j, h, w), flip in zip(boxes.tolist(), flips.tolist()):
        view = TF.resized_crop(img_chw, i, j, h, w, list(size), TF.InterpolationMode.BICUBIC, antialias=True)
        norm_i, norm_j = i / float(H), j / float(W)
        norm_h, norm_w = h / float(H), w / float(W)
        ys = torch.linspace(norm_i, norm_i + norm_h, hp)
        xs = torch.linspace(norm_j, norm_j + norm_w, wp)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        c = torch.stack([gx, gy], dim=-1)
        if flip:
            view = TF.hflip(view)
            c[:, :, 0] = (c[:, :, 0].max() - c[:, :, 0]) + c[:, :, 0].min()
        views.append(view)
        coords.append(c)
    return torch.stack(views), torch.stack(coords)
