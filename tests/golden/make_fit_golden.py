"""Pins oracle/fit.py against the REFERENCE's own code and writes tests/golden/fit_*.npz.

Runs only in the build container (needs /root/reference).  It imports, unmodified:
  * dvt/models/offline_denoiser.py::SingleImageDenoiser  (loaded by path, bypassing dvt/models/__init__.py which needs
    timm / tinycudann)
  * dvt/utils/misc.py::adjust_learning_rate
and drives them with the exact loop of main_img_denoising.py:39-89,121-130 (torch.optim.Adam, zero_grad, scaled
backward, step).  The only stand-in is the tiny-cuda-nn encoding, replaced by oracle/hashgrid.py inside a
NeuralFeatureField-shaped module (tinycudann is not installable here).  The script asserts that oracle/fit.py
reproduces the reference run, then stores small fixtures: inputs are re-derivable from seeds, outputs are stored.
"""
import argparse
import importlib.util
import os
import sys
import types
from itertools import chain

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import fit as OF  # noqa: E402
from oracle import hashgrid as HG  # noqa: E402


def load_reference():
    pkg = types.ModuleType("refdvt")
    pkg.__path__ = []
    sys.modules["refdvt"] = pkg
    nff = types.ModuleType("refdvt.neural_feature_field")

    class NeuralFeatureField(nn.Module):  # stand-in with the reference's structure (.neural_field params + .mlp)
        def __init__(self, feat_dim, meta):
            super().__init__()
            self.meta = meta
            self.table = nn.Parameter(torch.zeros(meta.n_params))
            self.mlp = nn.Sequential(nn.Linear(meta.n_output_dims, feat_dim // 2), nn.ReLU(),
                                     nn.Linear(feat_dim // 2, feat_dim))

        def forward(self, coords):
            assert coords.max() <= 1 and coords.min() >= 0, "coordinates should be in [0, 1]"
            enc = HG.encode(self.table, coords.view(-1, 2), self.meta)
            return self.mlp(enc.view(list(coords.shape[:-1]) + [-1]))

    nff.NeuralFeatureField = NeuralFeatureField
    sys.modules["refdvt.neural_feature_field"] = nff
    spec = importlib.util.spec_from_file_location("refdvt.offline_denoiser",
                                                  os.path.join(REF, "dvt/models/offline_denoiser.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "refdvt"
    sys.modules["refdvt.offline_denoiser"] = mod
    spec.loader.exec_module(mod)
    spec2 = importlib.util.spec_from_file_location("refdvt_misc", os.path.join(REF, "dvt/utils/misc.py"))
    misc = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(misc)
    return mod.SingleImageDenoiser, NeuralFeatureField, misc


def reference_run(cfg, feats, coords, meta, init, idx_stream):
    """main_img_denoising.py:39-89 + :121-130 with the reference classes."""
    SingleImageDenoiser, NeuralFeatureField, misc = load_reference()
    C, h, w = cfg["C"], cfg["h"], cfg["w"]
    denoiser = SingleImageDenoiser(noise_map_height=h, noise_map_width=w, feat_dim=C, layer_index=11)
    field = NeuralFeatureField(C, meta)
    with torch.no_grad():
        denoiser.shared_artifacts.copy_(init["G"])
        for i in (0, 2, 4):
            denoiser.residual_predictor[i].weight.copy_(init[f"res.{i}.weight"])
            denoiser.residual_predictor[i].bias.copy_(init[f"res.{i}.bias"])
        field.table.copy_(init["table"])
        for i in (0, 2):
            field.mlp[i].weight.copy_(init[f"mlp.{i}.weight"])
            field.mlp[i].bias.copy_(init[f"mlp.{i}.bias"])
    args = argparse.Namespace(lr=cfg["lr"], min_lr=cfg["min_lr"], warmup_iters=cfg["warmup_iters"],
                              num_iters=cfg["num_iters"])
    optimizer = torch.optim.Adam(chain(denoiser.parameters(), field.parameters()), lr=args.lr, eps=1e-15,
                                 weight_decay=cfg["weight_decay"], betas=(0.9, 0.99))
    V = feats.shape[0]
    sac = OF.make_patch_coordinates(h, w).unsqueeze(0).repeat(V, 1, 1, 1).reshape(-1, 2)
    braw = feats.reshape(-1, C)
    bpix = coords.reshape(-1, 2)
    logs = []
    for step in range(args.num_iters):
        denoiser.train()
        field.train()
        if step > int(cfg["freeze_after"] * args.num_iters):
            denoiser.stop_shared_artifacts_grad()
            denoiser.start_residual_predictor()
        idx = idx_stream[step]
        misc.adjust_learning_rate(optimizer, step, args)
        out = denoiser(raw_vit_outputs=braw[idx], global_pixel_coords=bpix[idx], neural_field=field,
                       shared_artifact_coords=sac[idx], return_visualization=False)
        loss = out["loss"]
        optimizer.zero_grad()
        (loss * cfg["loss_scale"]).backward()   # grad_scaler.scale(loss).backward(); never unscaled (:88-89)
        optimizer.step()
        if step % cfg["log_every"] == 0 or step == args.num_iters - 1:
            logs.append([step] + [float(out[k].detach()) if k in out else 0.0 for k in
                                  ("loss", "patch_l2_loss", "cosine_similarity_loss", "residual_loss",
                                   "residual_sparsity_loss")])
    with torch.no_grad():
        q = denoiser(raw_vit_outputs=feats[-1:], global_pixel_coords=coords[-1:], neural_field=field,
                     return_visualization=True)
    return {"logs": np.array(logs), "denoised_feats": q["denoised_feats"], "denoised_features": q["denoised_features"],
            "G": denoiser.shared_artifacts.detach(), "table": field.table.detach()}


CONFIGS = {
    # name: small enough for seconds on CPU; dims are multiples of 64 so the CUDA tiles are exercised with tails
    "small_L6": dict(C=128, h=8, w=8, V=6, bsz=256, n_levels=6, num_iters=300, warmup_iters=30, lr=0.01, min_lr=0.001,
                     weight_decay=1e-5, freeze_after=0.5, loss_scale=1024.0, log_every=10, seed=0),
    "small_L6_ls1": dict(C=128, h=8, w=8, V=6, bsz=256, n_levels=6, num_iters=120, warmup_iters=12, lr=0.01,
                         min_lr=0.001, weight_decay=1e-5, freeze_after=0.5, loss_scale=1.0, log_every=10, seed=1),
    "hashed_L16": dict(C=64, h=6, w=6, V=4, bsz=128, n_levels=16, num_iters=40, warmup_iters=4, lr=0.01, min_lr=0.001,
                       weight_decay=1e-5, freeze_after=0.5, loss_scale=1024.0, log_every=5, seed=2),
}


def make(name, cfg):
    torch.manual_seed(0)
    meta = HG.grid_meta(cfg["n_levels"])
    feats, coords = OF.synthetic_bank(cfg["V"], cfg["h"], cfg["w"], cfg["C"], seed=cfg["seed"])
    init = OF.init_params(cfg["C"], cfg["h"], cfg["w"], meta, seed=cfg["seed"])
    rs = np.random.RandomState(cfg["seed"])
    idx_stream = rs.randint(0, cfg["V"] * cfg["h"] * cfg["w"], (cfg["num_iters"], cfg["bsz"]))
    ref = reference_run(cfg, feats, coords, meta, init, idx_stream)
    ora = OF.fit(feats, coords, cfg["h"], cfg["w"], meta, init, idx_stream, lr=cfg["lr"], min_lr=cfg["min_lr"],
                 weight_decay=cfg["weight_decay"], warmup_iters=cfg["warmup_iters"], freeze_after=cfg["freeze_after"],
                 loss_scale=cfg["loss_scale"], log_every=cfg["log_every"])
    # --- the oracle must reproduce the reference run (same torch ops on the same machine -> tight) ---
    assert np.allclose(ref["logs"], ora["logs"], rtol=1e-4, atol=1e-6), np.abs(ref["logs"] - ora["logs"]).max()
    d = (ref["denoised_feats"] - ora["denoised_feats"]).abs().max().item()
    assert d < 1e-4, d
    assert (ref["G"] - ora["params"]["G"]).abs().max().item() < 1e-5
    assert (ref["table"] - ora["params"]["table"]).abs().max().item() < 1e-6
    arrays = {"cfg_keys": np.array(sorted(cfg)), "cfg_vals": np.array([float(cfg[k]) for k in sorted(cfg)]),
              "logs": ref["logs"], "denoised_feats": ref["denoised_feats"].numpy().astype(np.float32),
              "denoised_features": ref["denoised_features"].numpy().astype(np.float32),
              "G_final": ref["G"].numpy().astype(np.float32),
              "idx_checksum": np.array([int(idx_stream.sum())]),
              "table_sum": np.array([float(ref["table"].double().sum()), float(ref["table"].double().abs().sum())])}
    path = os.path.join(HERE, f"fit_{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: reference == oracle (max |d denoised| {d:.2e}); final loss {ref['logs'][-1][1]:.4f}; wrote {path} "
          f"{os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    for n, c in CONFIGS.items():
        make(n, c)
