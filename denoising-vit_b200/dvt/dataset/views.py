"""Host-side view generation for stage 1 (the step in front of HP-1).  Same classes and outputs as the reference's
dvt/dataset/single_image_dataset.py:12-51 and dvt/dataset/transform.py:9-76: one image -> `num_views` random resized
crops (bicubic, antialias, optional horizontal flip) plus, per view, the [h, w, 2] grid of (x, y) patch coordinates
of the crop inside the original image, normalised to [0, 1]."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torchvision
import torchvision.transforms.functional as TF
from PIL import Image


class RandomResizedCropFlip(torchvision.transforms.RandomResizedCrop):
    def __init__(self, size, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                 interpolation=TF.InterpolationMode.BICUBIC, antialias: Optional[bool] = True, horizontal_flip=True,
                 patch_size: Optional[int] = 14, stride: Optional[int] = 14):
        super().__init__(size, scale=scale, ratio=ratio, interpolation=interpolation, antialias=antialias)
        self.patch_size, self.stride, self.horizontal_flip = patch_size, stride, horizontal_flip
        self.h_patches = (self.size[0] - patch_size) // stride + 1
        self.w_patches = (self.size[1] - patch_size) // stride + 1

    def forward(self, img):
        top, left, ch, cw = self.get_params(img, self.scale, self.ratio)
        _, H, W = TF.get_dimensions(img)
        view = TF.resized_crop(img, top, left, ch, cw, self.size, self.interpolation, antialias=self.antialias)
        ys = torch.linspace(top / float(H), top / float(H) + ch / float(H), self.h_patches)
        xs = torch.linspace(left / float(W), left / float(W) + cw / float(W), self.w_patches)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        coords = torch.stack([gx, gy], dim=-1)
        if self.horizontal_flip and np.random.random() < 0.5:
            view = TF.hflip(view)
            x = coords[:, :, 0]
            coords[:, :, 0] = (x.max() - x) + x.min()
        return view, coords


class SingleImageDataset(torch.utils.data.Dataset):
    """Yields `num_views` augmented views of one image (set with `set_image`)."""

    def __init__(self, size, base_transform, final_transform, num_views: int = 768):
        self.size, self.base_transform, self.final_transform, self.num_views = size, base_transform, final_transform, num_views
        self.image = None

    def set_image(self, path: str):
        img = Image.open(path).convert("RGB")
        self.image = self.base_transform(np.array(img))

    def __len__(self):
        return self.num_views

    def __getitem__(self, i):
        view, coords = self.final_transform(self.image)
        return {"transformed_view": view, "pixel_coords": coords, "full_image": self.image}
