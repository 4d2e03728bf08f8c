#!/usr/bin/env python
"""How far does the REFERENCE's 2000-step fit at the headline size move when only the floating-point summation order
changes?  Same inputs as tests/golden/fit_headline_2000.npz, but the 2048 sampled rows of every step are visited in a
different order (a permutation inside each step: mathematically the identical batch, loss and gradient).  The distance of
the resulting `denoised_feats` from the golden is the noise floor any other implementation has to be judged against.
Runs the oracle restatement (bit-identical to the reference classes, tests/golden/make_fit_golden.py).  ~12 min on 8 cores."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_fit_golden_headline as H  # noqa: E402
from oracle import fit as OF  # noqa: E402

if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("DVT_GOLDEN_THREADS", os.cpu_count())))
    cfg = dict(H.CFG)
    if len(sys.argv) > 1:
        cfg["num_iters"] = int(sys.argv[1])
    meta, feats, coords, init, idx = H.inputs(cfg)
    rs = np.random.RandomState(12345)
    idx_perm = np.stack([row[rs.permutation(row.shape[0])] for row in idx])
    out = OF.fit(feats, coords, cfg["h"], cfg["w"], meta, init, idx_perm, lr=cfg["lr"], min_lr=cfg["min_lr"],
                 weight_decay=cfg["weight_decay"], warmup_iters=cfg["warmup_iters"], freeze_after=cfg["freeze_after"],
                 loss_scale=cfg["loss_scale"], log_every=cfg["log_every"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "fit_headline_2000.npz"))
    if cfg["num_iters"] == H.CFG["num_iters"]:
        ref = torch.from_numpy(z["denoised_feats"].astype(np.float32))
        got = out["denoised_feats"]
        cos = F.cosine_similarity(got.reshape(-1, cfg["C"]), ref.reshape(-1, cfg["C"]), dim=-1)
        print(f"permuted-batch oracle vs golden: min cosine {cos.min().item():.6f}  mean {cos.mean().item():.6f}  "
              f"1% quantile {cos.quantile(0.01).item():.6f}")
        dl = np.abs(out["logs"][:, 1:] - z["logs"][:, 1:]) / (np.abs(z["logs"][:, 1:]) + 1e-3)
        print(f"worst relative loss deviation over the logged steps: {dl.max():.4f}")
        np.savez_compressed("/tmp/noise_floor.npz", cos=cos.numpy(), logs=out["logs"])
