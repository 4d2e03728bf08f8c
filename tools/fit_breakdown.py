#!/usr/bin/env python
"""Times the HP-2 fit at the headline size (C=768, 37x37, 769 views, 2048 pixels, 16 levels): steps/s in phase 1 and 2,
with CUDA graphs on/off.  Also the target of ncu launch lists (--iters small)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
import dvt.models as DVT  # noqa: E402
from dvt.fit import FitEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--graph-steps", type=int, default=10)
    ap.add_argument("--views", type=int, default=769)
    ap.add_argument("--configs", default="",
                    help="';'-separated engine configurations 'pipeline:sweep_ctas[:graph_steps[:pdl]]', e.g. '1:48,32;0:0:20:0'")
    ap.add_argument("--graphs-only", action="store_true")
    a = ap.parse_args()
    C, h, w, V, bsz = 768, 37, 37, a.views, 2048
    field = DVT.NeuralFeatureField(feat_dim=C, n_levels=16)
    den = DVT.SingleImageDenoiser(h, w, C)
    g = torch.Generator(device="cuda").manual_seed(0)
    bank = torch.randn(V * h * w, C, device="cuda", generator=g)
    coords = torch.rand(V * h * w, 2, device="cuda", generator=g)
    idx = np.random.RandomState(0).randint(0, V * h * w, (a.iters, bsz))
    for cfg in (a.configs.split(";") if a.configs else [""]):
        gsteps = a.graph_steps
        if cfg:
            parts = cfg.split(":")
            os.environ["DVT_FIT_PIPELINE"] = parts[0]
            os.environ["DVT_FIT_SWEEP_CTAS"] = parts[1]
            if len(parts) > 2:
                gsteps = int(parts[2])
            if len(parts) > 3:
                os.environ["DVT_FIT_PDL"] = parts[3]
            print(f"== pipeline={parts[0]} sweep_ctas={parts[1]} graph_steps={gsteps} pdl={os.environ.get('DVT_FIT_PDL', 'default')}", flush=True)
        eng = FitEngine(C, h, w, bsz, field.meta)
        eng.load_modules(den, field)
        time_engine(eng, a, bank, coords, idx, (gsteps,) if a.graphs_only else (gsteps, 0))
        del eng


def time_engine(eng, a, bank, coords, idx, graph_modes):
    hyper = dict(lr=0.01, min_lr=0.001, warmup_iters=a.iters // 10, freeze_after=0.5, weight_decay=1e-5, loss_scale=1024.0)
    for gs in graph_modes:
        eng.begin(bank, coords, idx, **hyper)   # warm-up pass: graph capture / instantiation, first-touch, clocks
        eng.run(a.iters, graph_steps=gs)
        torch.cuda.synchronize()
        eng.begin(bank, coords, idx, **hyper)
        half = a.iters // 2 + 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        eng.run(half, graph_steps=gs)
        ev[1].record()
        eng.run(a.iters - half, graph_steps=gs)
        ev[2].record()
        torch.cuda.synchronize()
        p1, p2 = ev[0].elapsed_time(ev[1]) / half, ev[1].elapsed_time(ev[2]) / (a.iters - half)
        print(f"graph_steps={gs:3d}: phase1 {p1 * 1e3:8.1f} us/step   phase2 {p2 * 1e3:8.1f} us/step")


if __name__ == "__main__":
    main()
