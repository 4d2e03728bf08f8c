"""Drop-in for dvt/models/offline_denoiser.py: `SingleImageDenoiser` with the reference's constructor, parameters
(`shared_artifacts [1,C,h,w]`, `residual_predictor.{0,2,4}`), phase switches and `forward` contract
(reference offline_denoiser.py:11-171).  The module forward is the generic (torch-op) statement of the model used
for inspection / visualisation; the per-image optimisation loop runs in the fused CUDA engine (dvt/fit.py), which
reads and writes this module's parameters."""
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

    def forward(self, raw_vit_outputs: Tensor, global_pixel_coords: Tensor, neural_field: NeuralFeatureField = None,
                shared_artifact_coords: Tensor = None, return_visualization: bool = False) -> Dict[str, Tensor]:
        C = self.feat_dim
        if raw_vit_outputs.dim() != 2:
            original_shape = raw_vit_outputs.shape
            raw = raw_vit_outputs.reshape(-1, C)
            coords = global_pixel_coords.reshape(-1, 2)
            shared = self.shared_artifacts.permute(0, 2, 3, 1).reshape(-1, C)
        else:
            assert shared_artifact_coords is not None, "shared_artifact_coords must be provided."
            original_shape = None
            raw, coords = raw_vit_outputs, global_pixel_coords
            shared = F.grid_sample(self.shared_artifacts, shared_artifact_coords[None, None, ...], mode="bilinear",
                                   align_corners=True)
            shared = shared.reshape(C, -1).permute(1, 0)
        denoised = neural_field(coords)
        pred_residual = self.residual_predictor(raw) if self.use_residual_predictor else None
        pred = denoised + shared + pred_residual.detach() if pred_residual is not None else shared + denoised
        patch_l2_loss = F.mse_loss(pred, raw)
        cosine_similarity_loss = 1 - F.cosine_similarity(pred, raw, dim=-1).mean()
        loss = patch_l2_loss + cosine_similarity_loss
        results = {"patch_l2_loss": patch_l2_loss, "loss": loss, "cosine_similarity_loss": cosine_similarity_loss}
        if pred_residual is not None:
            gt_residual = (raw - denoised - shared).detach()
            residual_loss = 0.1 * F.mse_loss(pred_residual, gt_residual)
            residual_sparsity_loss = 0.02 * pred_residual.abs().mean()
            loss = loss + residual_loss + residual_sparsity_loss
            results.update(loss=loss, residual_loss=residual_loss, residual_sparsity_loss=residual_sparsity_loss)
        if return_visualization:
            assert original_shape is not None, "original_shape must be provided."
            sh = original_shape[:-1]
            results["raw_vit_outputs"] = raw.detach().reshape(*sh, -1)
            results["pred_features"] = pred.detach().reshape(*sh, -1)
            results["denoised_feats"] = denoised.detach().reshape(*sh, -1)
            results["shared_patterns"] = shared.detach().reshape(*sh, -1)
            if pred_residual is not None:
                results["pred_residual"] = pred_residual.detach().reshape(*sh, -1)
                results["shared_patterns_and_residual"] = (shared + pred_residual).detach().reshape(*sh, -1)
                denoised_features = raw - shared - pred_residual
            else:
                denoised_features = raw - shared
            results["denoised_features"] = denoised_features.detach().reshape(*sh, -1)
        return results
