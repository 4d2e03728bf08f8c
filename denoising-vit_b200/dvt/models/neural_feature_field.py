"""Drop-in for dvt/models/neural_feature_field.py: same constructor and `forward(coords in [0,1]) -> features`, with
the tiny-cuda-nn HashGrid replaced by the hash-grid kernels of libdvt_b200.so.  The per-image fit does not go
through this module's autograd (it runs the fused engine in dvt/fit.py); the module forward/backward exists so the
object remains a normal trainable nn.Module with the reference's parameter structure:
    neural_field.neural_field.params  (flat fp32 table, like tcnn.Encoding.params)
    neural_field.mlp.{0,2}.{weight,bias}
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib
from .._lib import check, cur_stream, lib, ptr
from .hashgrid_meta import HashGridMeta, make_meta


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords: Tensor, params: Tensor, meta: HashGridMeta):
        if not coords.is_cuda:
            raise _lib.DvtError("dvt_b200 hash grid needs CUDA tensors (no CPU fallback)")
        coords = coords.contiguous().float()
        n = coords.shape[0]
        out = torch.empty((n, meta.n_output_dims), device=coords.device, dtype=torch.float32)
        check(lib().dvt_hashgrid_fwd(*meta.c_args(), ptr(params), ptr(coords), n, ptr(out), cur_stream()),
              "dvt_hashgrid_fwd")
        ctx.save_for_backward(coords)
        ctx.meta = meta
        ctx.n_params = params.numel()
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        (coords,) = ctx.saved_tensors
        meta = ctx.meta
        g = torch.zeros(ctx.n_params, device=dout.device, dtype=torch.float32)  # dense gradient, like tcnn
        dout = dout.contiguous().float()
        check(lib().dvt_hashgrid_bwd(*meta.c_args(), ptr(coords), coords.shape[0], ptr(dout), ptr(g), cur_stream()),
              "dvt_hashgrid_bwd")
        return None, g, None


class HashGridEncoding(nn.Module):
    """Stands where `tcnn.Encoding(n_input_dims=2, HashGrid{...}, dtype=float32)` stood
    (reference neural_feature_field.py:25-39): `.params` flat fp32, `.n_output_dims`."""

    def __init__(self, meta: HashGridMeta, seed: int = 1337):
        super().__init__()
        self.meta = meta
        self.n_input_dims = 2
        self.n_output_dims = meta.n_output_dims
        g = torch.Generator().manual_seed(seed)
        # tcnn initialises grid parameters U(-1e-4, 1e-4) (its own pcg32 stream; not reproducible without tcnn)
        self.params = nn.Parameter((torch.rand(meta.n_params, generator=g) * 2 - 1) * 1e-4)

    def forward(self, coords: Tensor) -> Tensor:
        return _HashGridFn.apply(coords, self.params, self.meta)


class NeuralFeatureField(nn.Module):
    """A neural field that maps 2D coordinates to features."""

    def __init__(self, feat_dim: int = 768, base_resolution: int = 16, max_resolution: int = 1024, n_levels: int = 10,
                 n_features_per_level: int = 8, log2_hashmap_size: int = 20):
        super().__init__()
        self.meta = make_meta(n_levels, base_resolution, max_resolution, n_features_per_level, log2_hashmap_size)
        self.neural_field = HashGridEncoding(self.meta)
        self.mlp = nn.Sequential(
            nn.Linear(self.neural_field.n_output_dims, feat_dim // 2),
            nn.ReLU(),
            nn.Linear(feat_dim // 2, feat_dim),
        )

    def forward(self, coords: Tensor):
        assert coords.max() <= 1 and coords.min() >= 0, "coordinates should be in [0, 1]"
        denoised_features = self.neural_field(coords.reshape(-1, 2))
        return self.mlp(denoised_features.view(list(coords.shape[:-1]) + [-1]))
