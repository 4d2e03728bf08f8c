#!/usr/bin/env python
"""Times HP-1 and HP-2 separately on (a) white-noise views and (b) views generated from one image (8(f-1)), to see
whether the paths are data dependent (attention rescaling, hash-grid contention)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
sys.path.insert(0, ROOT)
import dvt.models as DVT  # noqa: E402
from dvt.dataset import GpuViewGenerator  # noqa: E402
from dvt.stage1 import Stage1Config, Stage1Pipeline  # noqa: E402
import bench  # noqa: E402


def t(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = "cuda"
    V = 769
    vit = DVT.PretrainedViTWrapper(bench.MODEL, stride=14, allow_random_init=True)
    with torch.no_grad():
        for b in vit.model.blocks:
            b.ls1.gamma.fill_(1.0)
            b.ls2.gamma.fill_(1.0)
    vit = vit.to(dev).eval()
    cfg = Stage1Config(num_iters=2000, warmup_iters=200, n_levels=16, extract_bsz=32, pixel_bsz=2048, graph_steps=20)
    pipe = Stage1Pipeline(vit, 11, (518, 518), cfg)
    noise = torch.randn(V, 3, 518, 518, device=dev)
    coords_syn = bench.synthetic_coords(V, pipe.h, pipe.w, 0, dev)
    gen = GpuViewGenerator((518, 518), num_views=768)
    img = torch.randn(3, 518, 518, device=dev)
    views, coords_gen = gen(img)
    smooth = torch.nn.functional.avg_pool2d(torch.randn(1, 3, 518, 518, device=dev), 31, 1, 15)[0].contiguous() * 8
    views_s, coords_s = gen(smooth)
    idx = np.random.RandomState(0).randint(0, V * pipe.h * pipe.w, (2000, 2048))
    for name, vw, co in (("white-noise views", noise, coords_syn), ("crops of a white-noise image", views, coords_gen),
                         ("crops of a smooth image", views_s, coords_s)):
        ms1 = t(lambda: pipe.extract_bank(vw))
        bank = pipe.extract_bank(vw)
        fin = bool(torch.isfinite(bank).all())
        ms2 = t(lambda: pipe.denoise(bank, co, idx), reps=1)
        print(f"{name:32s} HP-1 {ms1:7.1f} ms   HP-2 {ms2:7.1f} ms   bank finite {fin}  bank |max| {bank.abs().max().item():.3g}")
    print(f"view generation: {t(lambda: gen(img)):.1f} ms per image (host sampling + kernel)")


if __name__ == "__main__":
    main()
