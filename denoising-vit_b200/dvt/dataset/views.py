"""Image loading for stage 1.  The reference decodes the image, resizes it to the working resolution and normalises it
once per image (`SingleImageDataset.set_image` + its `base_transform`, dvt/dataset/single_image_dataset.py:26-27,
main_img_denoising.py:277-287); every augmented view is then cut from that tensor.  Here the views are cut on the GPU
(dvt/dataset/gpu_views.py), so the host side is only this loader -- there is no CPU view pipeline."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torchvision.transforms.functional as TF
from PIL import Image


def load_image(path: str, size: Tuple[int, int], mean: Sequence[float], std: Sequence[float]) -> torch.Tensor:
    """RGB image file -> normalised float32 [3, H, W] at `size` (PIL bilinear resize with antialiasing, like
    `transforms.Resize(size)` on a PIL image, then ToTensor and Normalize)."""
    with Image.open(path) as im:
        im = im.convert("RGB")
        im = TF.resize(im, [int(size[0]), int(size[1])])
        x = TF.to_tensor(im)
    return TF.normalize(x, list(mean), list(std))
