#!/usr/bin/env python
"""Times GPU view generation (f-1): 768 random-resized-crop(+flip) views of a 518^2 image to [768, 3, 518, 518], fp32 and bf16
(`dvt_view_crops`).  CUDA events, 10 calls after 3 warm-ups; prints ms per image and the output bandwidth."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
from dvt.dataset import GpuViewGenerator  # noqa: E402

torch.manual_seed(0)
np.random.seed(0)
img = torch.rand(3, 518, 518, device="cuda")
for dt in (torch.float32, torch.bfloat16):
    gen = GpuViewGenerator((518, 518), num_views=768, dtype=dt, append_full_image=False)
    out = torch.empty((768, 3, 518, 518), device="cuda", dtype=dt)
    for _ in range(3):
        gen(img, views_out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gen(img, views_out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"view_crops {str(dt):16s} {ms:7.3f} ms / image (768 views, host box sampling included)   "
          f"{out.numel() * out.element_size() / ms / 1e6:8.1f} GB/s written")
