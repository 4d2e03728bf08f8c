"""GPU view generation for stage 1: the 768 random-resized-crop(+flip) views of an image and
their patch-coordinate grids are produced by ONE kernel launch from the device-resident image (`dvt_view_crops`,
csrc/views.cu) instead of 768 CPU `resized_crop` calls in DataLoader workers (reference:
dvt/dataset/transform.py:39-76, dvt/dataset/single_image_dataset.py:29-48, main_img_denoising.py:277-310).

Only the parameters are drawn on the host, with the reference's own RNG calls in the reference's order
(`RandomResizedCrop.get_params` on the torch RNG, then `np.random.random() < 0.5` for the flip), so a seeded run draws
the boxes a single-process run of the reference transform would draw."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torchvision

from .. import ops


def sample_view_params(img_chw: torch.Tensor, num_views: int, scale=(0.1, 0.5), ratio=(3.0 / 4.0, 4.0 / 3.0),
                       horizontal_flip: bool = True, flip_rng=None) -> Tuple[np.ndarray, np.ndarray]:
    """boxes int32 [V, 4] = (top, left, height, width), flips int32 [V] (transform.py:48 and :69, per view).
    flip_rng: where the flip decisions are drawn (default: the global numpy RNG, like a single-process run of the
    reference transform).  The reference's stage-1 driver runs the transform in DataLoader workers, each with its own
    re-seeded numpy RNG, so its MAIN process's RNG is consumed by the sampling stream only; a driver that wants that
    stream to stay the reference's passes a private `np.random.RandomState` here."""
    rng = np.random if flip_rng is None else flip_rng
    probe = torch.empty((3, int(img_chw.shape[-2]), int(img_chw.shape[-1])), device="meta")  # get_params reads the size only
    boxes = np.empty((num_views, 4), dtype=np.int32)
    flips = np.empty((num_views,), dtype=np.int32)
    for v in range(num_views):
        boxes[v] = torchvision.transforms.RandomResizedCrop.get_params(probe, list(scale), list(ratio))
        flips[v] = int(horizontal_flip and rng.random_sample() < 0.5)
    return boxes, flips


class GpuViewGenerator:
    """Drop-in for `RandomResizedCropFlip` + `SingleImageDataset` + the collecting DataLoader loop: all views of one image
    at once.  `__call__(image)` returns (views [V(+1), 3, H, W], coords [V(+1), hp, wp, 2]); with `append_full_image` the
    un-augmented image and its linspace(0, 1) grid are appended as the last view (main_img_denoising.py:325-337)."""

    def __init__(self, size, num_views: int = 768, scale=(0.1, 0.5), ratio=(3.0 / 4.0, 4.0 / 3.0), patch_size: int = 14,
                 stride: int = 14, horizontal_flip: bool = True, dtype: torch.dtype = torch.float32,
                 append_full_image: bool = True, flip_rng=None):
        self.size = (int(size[0]), int(size[1]))
        self.num_views, self.scale, self.ratio = num_views, tuple(scale), tuple(ratio)
        self.horizontal_flip, self.dtype, self.append_full_image = horizontal_flip, dtype, append_full_image
        self.h_patches = (self.size[0] - patch_size) // stride + 1
        self.w_patches = (self.size[1] - patch_size) // stride + 1
        self._full_grid = None
        self.flip_rng = flip_rng

    def __call__(self, image: torch.Tensor, views_out: Optional[torch.Tensor] = None,
                 coords_out: Optional[torch.Tensor] = None):
        if not image.is_cuda:
            raise ops.DvtError("GpuViewGenerator needs the image on the GPU (no CPU fallback)")
        assert image.dim() == 3 and image.shape[0] == 3
        V, extra = self.num_views, int(self.append_full_image)
        OH, OW = self.size
        if extra:
            assert tuple(image.shape[-2:]) == self.size, "appending the full image needs it at the view size"
        if views_out is None:
            views_out = torch.empty((V + extra, 3, OH, OW), device=image.device, dtype=self.dtype)
        if coords_out is None:
            coords_out = torch.empty((V + extra, self.h_patches, self.w_patches, 2), device=image.device, dtype=torch.float32)
        boxes, flips = sample_view_params(image, V, self.scale, self.ratio, self.horizontal_flip, self.flip_rng)
        ops.view_crops(image, boxes, flips, self.size, self.h_patches, self.w_patches, views_out[:V], coords_out[:V])
        if extra:
            views_out[V].copy_(image)
            if self._full_grid is None or self._full_grid.device != image.device:
                ys, xs = torch.linspace(0, 1, self.h_patches), torch.linspace(0, 1, self.w_patches)
                gy, gx = torch.meshgrid(ys, xs, indexing="ij")
                self._full_grid = torch.stack([gx, gy], dim=-1).to(image.device)   # uploaded once: no per-image host sync
            coords_out[V].copy_(self._full_grid)
        self.last_boxes, self.last_flips = boxes, flips
        return views_out, coords_out
