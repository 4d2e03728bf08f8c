"""Drop-in for dvt/models/offline_denoiser.py: `SingleImageDenoiser` with the reference's constructor, parameters
(`shared_artifacts [1,C,h,w]`, `residual_predictor.{0,2,4}`), phase switches and `forward` contract
(reference offline_denoiser.py:11-171).  The module forward is the generic (torch-op) statement of the model used
for inspection / visualisation; the per-image optimisation loop runs in the fused CUDA engine (dvt/fit.py), which
reads and writes this module's parameters.  `forward` is written from the model's definition in terms of its own
pieces (`artifact_rows`, `loss_terms`); tests/test_oracle_fit.py checks it against the oracle restatement."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .neural_feature_field import NeuralFeatureField


class SingleImageDenoiser(nn.Module):
    def __init__(self, noise_map_height: int = 37, noise_map_width: int = 37, feat_dim: int = 768,
                 layer_index: int = 11, enable_residual_predictor: bool = True, disable_pe: bool = False):
        super().__init__()
        self.noise_map_h = noise_map_height
        self.noise_map_w = noise_map_width
        self.feat_dim = feat_dim
        self.layer_idx = layer_index
        if disable_pe:
            self.shared_artifacts = nn.Parameter(torch.zeros(1, feat_dim, noise_map_height, noise_map_width),
                                                 requires_grad=False)
        else:
            self.shared_artifacts = nn.Parameter(torch.randn(1, feat_dim, noise_map_height, noise_map_width) * 0.02,
                                                 requires_grad=True)
        self.enable_residual_predictor = enable_residual_predictor
        if self.enable_residual_predictor:
            self.residual_predictor = nn.Sequential(
                nn.Linear(feat_dim, feat_dim // 4), nn.ReLU(),
                nn.Linear(feat_dim // 4, feat_dim // 4), nn.ReLU(),
                nn.Linear(feat_dim // 4, feat_dim),
            )
        self.residual_predictor_start = False

    def start_residual_predictor(self):
        self.residual_predictor_start = True

    @property
    def use_residual_predictor(self):
        return self.enable_residual_predictor and self.residual_predictor_start

    def stop_shared_artifacts_grad(self):
        self.shared_artifacts.requires_grad = False

    # ---- pieces of the model, each usable on its own -----------------------------------------------------------
    def artifact_rows(self, shared_artifact_coords: Tensor = None) -> Tensor:
        """Rows of the shared artifact map G as [n, C].  Without coordinates: all h*w cells in raster order (the 4-D
        "visualisation" call).  With coordinates in [-1, 1] (x, y): bilinear interpolation between the four surrounding
        cells with `align_corners=True` semantics; at exact grid nodes -- the only place the stage-1 loop samples,
        main_img_denoising.py:58-62 -- this is G[:, r, c] (what the CUDA engine indexes directly, fit.cu `cell`)."""
        h, w, C = self.noise_map_h, self.noise_map_w, self.feat_dim
        cells = self.shared_artifacts[0].reshape(C, h * w).t()          # [h*w, C] view of the parameter
        if shared_artifact_coords is None:
            return cells
        fx = (shared_artifact_coords[:, 0].float() + 1) * (0.5 * (w - 1))
        fy = (shared_artifact_coords[:, 1].float() + 1) * (0.5 * (h - 1))
        x0, y0 = fx.floor(), fy.floor()
        out = 0
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi = x0 + dx, y0 + dy
                wgt = (1 - (fx - xi).abs()) * (1 - (fy - yi).abs())
                inside = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)     # zero padding outside the map
                idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long()
                out = out + cells[idx] * (wgt * inside).unsqueeze(-1)
        return out

    @staticmethod
    def loss_terms(pred: Tensor, raw: Tensor, pred_residual: Tensor = None, residual_target: Tensor = None):
        """The reference's objective (offline_denoiser.py:122-138): MSE + (1 - mean cosine) on the prediction, plus -- once
        the residual predictor runs -- 0.1 MSE(residual, target) + 0.02 mean |residual|.  The CUDA engine evaluates the same
        terms and their gradients in `fit_loss_kernel`."""
        terms = {"patch_l2_loss": (pred - raw).square().mean(),
                 "cosine_similarity_loss": 1 - F.cosine_similarity(pred, raw, dim=-1).mean()}
        total = terms["patch_l2_loss"] + terms["cosine_similarity_loss"]
        if pred_residual is not None:
            terms["residual_loss"] = 0.1 * (pred_residual - residual_target).square().mean()
            terms["residual_sparsity_loss"] = 0.02 * pred_residual.abs().mean()
            total = total + terms["residual_loss"] + terms["residual_sparsity_loss"]
        terms["loss"] = total
        return terms

    def forward(self, raw_vit_outputs: Tensor, global_pixel_coords: Tensor, neural_field: NeuralFeatureField = None,
                shared_artifact_coords: Tensor = None, return_visualization: bool = False) -> Dict[str, Tensor]:
        """Same contract as the reference (offline_denoiser.py:62-171).  2-D inputs [n, C] + per-row artifact coordinates:
        the training call; any other rank: one whole map per leading index, G used cell by cell."""
        C = self.feat_dim
        per_row = raw_vit_outputs.dim() == 2
        if per_row and shared_artifact_coords is None:
            raise AssertionError("shared_artifact_coords must be provided.")
        lead = None if per_row else tuple(raw_vit_outputs.shape[:-1])
        raw = raw_vit_outputs.reshape(-1, C)
        shared = self.artifact_rows(shared_artifact_coords if per_row else None)
        denoised = neural_field(global_pixel_coords.reshape(-1, 2))
        if self.use_residual_predictor:
            residual = self.residual_predictor(raw)
            pred = denoised + shared + residual.detach()
            results = self.loss_terms(pred, raw, residual, (raw - denoised - shared).detach())
        else:
            residual = None
            pred = shared + denoised
            results = self.loss_terms(pred, raw)
        if not return_visualization:
            return results
        if lead is None:
            raise AssertionError("original_shape must be provided.")
        maps = {"raw_vit_outputs": raw, "pred_features": pred, "denoised_feats": denoised, "shared_patterns": shared,
                "denoised_features": raw - shared if residual is None else raw - shared - residual}
        if residual is not None:
            maps["pred_residual"] = residual
            maps["shared_patterns_and_residual"] = shared + residual
        for k, v in maps.items():
            results[k] = v.detach().reshape(*lead, -1)
        return results
