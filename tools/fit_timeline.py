#!/usr/bin/env python
"""Kernel timeline of the HP-2 fit in steady state (CUPTI through torch.profiler: start / duration / stream of every kernel
of a few graph launches), written as CSV for offline reading.  Timing under a tracer is NOT a bench number; the point is
where the streams idle and what the critical path of one step is.

  python tools/fit_timeline.py --out gpurun_out/fit_timeline.csv [--phase 1|2] [--graph-steps 20]"""
import argparse
import csv
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
import dvt.models as DVT  # noqa: E402
from dvt.fit import FitEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/fit_timeline.csv")
    ap.add_argument("--graph-steps", type=int, default=20)
    ap.add_argument("--views", type=int, default=769)
    a = ap.parse_args()
    C, h, w, V, bsz, iters = 768, 37, 37, a.views, 2048, 400
    field = DVT.NeuralFeatureField(feat_dim=C, n_levels=16)
    den = DVT.SingleImageDenoiser(h, w, C)
    g = torch.Generator(device="cuda").manual_seed(0)
    bank = torch.randn(V * h * w, C, device="cuda", generator=g)
    coords = torch.rand(V * h * w, 2, device="cuda", generator=g)
    idx = np.random.RandomState(0).randint(0, V * h * w, (iters, bsz))
    eng = FitEngine(C, h, w, bsz, field.meta)
    eng.load_modules(den, field)
    hyper = dict(lr=0.01, min_lr=0.001, warmup_iters=iters // 10, freeze_after=0.5, weight_decay=1e-5, loss_scale=1024.0)
    eng.begin(bank, coords, idx, **hyper)
    eng.run(iters, graph_steps=a.graph_steps)  # warm-up: graphs instantiated for both phases
    torch.cuda.synchronize()
    rows = []
    for phase in (1, 2):
        eng.begin(bank, coords, idx, **hyper)
        skip = 40 if phase == 1 else iters // 2 + 1 + 2 * a.graph_steps
        eng.run(skip, graph_steps=a.graph_steps)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.run(3 * a.graph_steps, graph_steps=a.graph_steps)
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        t0 = min(e.time_range.start for e in evs)
        for e in sorted(evs, key=lambda e: e.time_range.start):
            rows.append((phase, e.name[:90], e.device_resource_id, round(e.time_range.start - t0, 3),
                         round(e.time_range.end - e.time_range.start, 3)))
        eng.run(iters - skip - 3 * a.graph_steps, graph_steps=a.graph_steps)
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["phase", "kernel", "stream", "start_us", "dur_us"])
        wr.writerows(rows)
    print(f"{len(rows)} kernel records -> {a.out}")


if __name__ == "__main__":
    main()
