#!/usr/bin/env python
"""Times the view generation at the headline size: 768 views of 518^2 from one image."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
from dvt import ops  # noqa: E402
from dvt.dataset import GpuViewGenerator, sample_view_params  # noqa: E402

img = torch.randn(3, 518, 518, device="cuda")
torch.manual_seed(0)
np.random.seed(0)
t0 = time.perf_counter()
boxes, flips = sample_view_params(img, 768)
t_host = time.perf_counter() - t0
for dt in (torch.float32, torch.bfloat16):
    out = torch.empty(768, 3, 518, 518, device="cuda", dtype=dt)
    co = torch.empty(768, 37, 37, 2, device="cuda")
    for _ in range(2):
        ops.view_crops(img, boxes, flips, (518, 518), 37, 37, out, co)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.view_crops(img, boxes, flips, (518, 518), 37, 37, out, co)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"view_crops 768 x 3 x 518 x 518 -> {dt}: {ms:.3f} ms  ({out.numel() * out.element_size() / ms / 1e6:.0f} GB/s written)")
gen = GpuViewGenerator((518, 518), num_views=768)
torch.cuda.synchronize()
t0 = time.perf_counter()
v, c = gen(img)
torch.cuda.synchronize()
print(f"host: sample_view_params(768) {t_host * 1e3:.1f} ms; GpuViewGenerator() end to end {1e3 * (time.perf_counter() - t0):.1f} ms")
