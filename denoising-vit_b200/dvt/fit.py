"""Host side of HP-2: drives the fused per-image fit engine of libdvt_b200.so.

`FitEngine.fit(...)` is what `denoise_an_image` calls instead of the reference's Python loop
(main_img_denoising.py:67-89): it uploads the parameters of a `SingleImageDenoiser` and a `NeuralFeatureField`,
replays the sampling stream, runs every optimisation step on the GPU (CUDA graphs, no host sync inside the loop)
and writes the optimised parameters back into the modules."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_void_p
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, lib, ptr
from .models.hashgrid_meta import HashGridMeta

LOSS_KEYS = ("loss", "patch_l2_loss", "cosine_similarity_loss", "residual_loss", "residual_sparsity_loss")

# engine parameter name -> (module, attribute path)
_DENOISER_PARAMS = {"G": "shared_artifacts", "res.0.weight": "residual_predictor.0.weight",
                    "res.0.bias": "residual_predictor.0.bias", "res.2.weight": "residual_predictor.2.weight",
                    "res.2.bias": "residual_predictor.2.bias", "res.4.weight": "residual_predictor.4.weight",
                    "res.4.bias": "residual_predictor.4.bias"}
_FIELD_PARAMS = {"table": "neural_field.params", "mlp.0.weight": "mlp.0.weight", "mlp.0.bias": "mlp.0.bias",
                 "mlp.2.weight": "mlp.2.weight", "mlp.2.bias": "mlp.2.bias"}


def _get(module, path):
    obj = module
    for part in path.split("."):
        obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
    return obj


def artifact_axis_table(size: int):
    """What `F.grid_sample(G, coords, bilinear, align_corners=True)` does with the reference's node coordinates
    `torch.linspace(-1, 1, size)` (main_img_denoising.py:21-25,58-62) along one axis, in the same fp32 arithmetic as ATen
    (`grid_sampler_unnormalize`: ((x + 1) / 2) * (size - 1); corner weights (i0 + 1) - ix and ix - i0): per node the first
    cell it touches and the weights of that cell and of the next one.  For 37 nodes, 9 are not exact integers in fp32 and
    leak ~1e-6 of their value / gradient into a neighbour."""
    x = torch.linspace(-1, 1, size, dtype=torch.float32)
    ix = ((x + 1.0) / 2) * (size - 1)
    i0 = torch.floor(ix)
    w1 = ix - i0
    w0 = (i0 + 1) - ix
    return i0.to(torch.int32), w0, w1


class FitEngine:
    def __init__(self, feat_dim: int, noise_map_height: int, noise_map_width: int, pixel_bsz: int, meta: HashGridMeta,
                 exact_grid_sample: bool = True):
        self.C, self.h, self.w, self.bsz, self.meta = feat_dim, noise_map_height, noise_map_width, pixel_bsz, meta
        h = c_void_p()
        check(lib().dvt_fit_create(byref(h), feat_dim, noise_map_height, noise_map_width, pixel_bsz, *meta.c_args()),
              "dvt_fit_create")
        self._h = h
        import os
        if exact_grid_sample and os.environ.get("DVT_FIT_EXACT_GRID", "1") != "0":
            # reproduce grid_sample's fp32 behaviour at the node coordinates (see artifact_axis_table)
            xi, xw0, xw1 = artifact_axis_table(noise_map_width)
            yi, yw0, yw1 = artifact_axis_table(noise_map_height)
            i0 = np.ascontiguousarray(torch.cat([xi, yi]).numpy())
            w0 = np.ascontiguousarray(torch.cat([xw0, yw0]).numpy())
            w1 = np.ascontiguousarray(torch.cat([xw1, yw1]).numpy())
            check(lib().dvt_fit_set_artifact_grid(self._h, i0.ctypes.data_as(c_void_p), w0.ctypes.data_as(c_void_p),
                                                  w1.ctypes.data_as(c_void_p)), "dvt_fit_set_artifact_grid")
        self.num_iters = 0
        self._keep = None

    def __del__(self):
        try:
            if self._h is not None:
                lib().dvt_fit_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------------------------
    def set_param(self, name: str, t: torch.Tensor):
        t = t.detach().to(dtype=torch.float32).contiguous()
        check(lib().dvt_fit_set_param(self._h, name.encode(), ptr(t), t.numel(), cur_stream()),
              f"dvt_fit_set_param({name})")

    def init_params(self, seed: int):
        """Fresh parameters for the next fit, drawn on the GPU (what constructing new SingleImageDenoiser /
        NeuralFeatureField modules per image does in the reference, main_img_denoising.py:39-47); no host round trip."""
        check(lib().dvt_fit_init_params(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, cur_stream()), "dvt_fit_init_params")

    def get_param(self, name: str, like: torch.Tensor) -> torch.Tensor:
        out = torch.empty(like.shape, device="cuda", dtype=torch.float32)
        check(lib().dvt_fit_get_param(self._h, name.encode(), ptr(out), out.numel()), f"dvt_fit_get_param({name})")
        return out

    def load_modules(self, denoiser, neural_field):
        for k, path in _DENOISER_PARAMS.items():
            self.set_param(k, _get(denoiser, path))
        for k, path in _FIELD_PARAMS.items():
            self.set_param(k, _get(neural_field, path))

    def store_modules(self, denoiser, neural_field):
        with torch.no_grad():
            for k, path in _DENOISER_PARAMS.items():
                p = _get(denoiser, path)
                p.copy_(self.get_param(k, p).to(p.device))
            for k, path in _FIELD_PARAMS.items():
                p = _get(neural_field, path)
                p.copy_(self.get_param(k, p).to(p.device))

    # ---- optimisation ----------------------------------------------------------------------------------
    def begin(self, bank_feats: torch.Tensor, bank_coords: torch.Tensor, idx_stream: np.ndarray, *, lr: float,
              min_lr: float, warmup_iters: int, freeze_after: float, weight_decay: float, loss_scale: float = 1024.0,
              validate: bool = True):
        """bank_feats [rows, C] f32 cuda, bank_coords [rows, 2] f32 cuda, idx_stream int [num_iters, bsz].
        validate=True reports out-of-range coordinates / rows here (blocks until the device has checked them, like the
        reference's assert); validate=False only enqueues -- call `check()` at the next natural synchronisation point."""
        if not (bank_feats.is_cuda and bank_coords.is_cuda):
            raise _lib.DvtError("the fit engine needs the feature bank on the GPU (no CPU fallback)")
        assert bank_feats.dtype == torch.float32 and bank_feats.is_contiguous() and bank_feats.shape[1] == self.C
        assert bank_coords.dtype == torch.float32 and bank_coords.is_contiguous()
        idx = np.ascontiguousarray(idx_stream, dtype=np.int32)
        assert idx.ndim == 2 and idx.shape[1] == self.bsz
        self.num_iters = idx.shape[0]
        self._keep = (bank_feats, bank_coords)  # borrowed by the engine
        # int(args.freeze_shared_artifacts_after * args.num_iters), in double like the reference (main_img_denoising.py:70)
        freeze_step = int(freeze_after * self.num_iters)
        check(lib().dvt_fit_begin(self._h, ptr(bank_feats), ptr(bank_coords), bank_feats.shape[0],
                                  idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), self.num_iters, float(lr),
                                  float(min_lr), int(warmup_iters), freeze_step, float(weight_decay), float(loss_scale),
                                  int(validate), cur_stream()), "dvt_fit_begin")

    def check(self):
        """Raises if a fit begun with validate=False saw coordinates outside [0, 1] or rows outside the bank.  Blocks."""
        check(lib().dvt_fit_check(self._h), "dvt_fit_check")

    def run(self, count: Optional[int] = None, graph_steps: int = 10):
        check(lib().dvt_fit_run(self._h, self.num_iters if count is None else count, graph_steps, cur_stream()),
              "dvt_fit_run")

    def losses(self) -> np.ndarray:
        out = np.zeros((self.num_iters, 5), np.float32)
        check(lib().dvt_fit_losses(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), self.num_iters),
              "dvt_fit_losses")
        return out

    def losses_async(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Loss table copied asynchronously on the current stream into `out` (pinned host or device, [num_iters, 5] f32)."""
        if out is None:
            out = torch.empty((self.num_iters, 5), dtype=torch.float32).pin_memory()
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == self.num_iters * 5
        check(lib().dvt_fit_losses_async(self._h, c_void_p(out.data_ptr()), self.num_iters, cur_stream()),
              "dvt_fit_losses_async")
        return out

    def query(self, coords: torch.Tensor, assume_valid: bool = False) -> torch.Tensor:
        """neural_field(coords): coords [..., 2] in [0,1] -> [..., C] (the `denoised_feats` of the reference).
        assume_valid: the caller has already range-checked `coords` (the check reads the tensor back, which blocks the
        host until everything queued before it -- e.g. a whole fit -- has finished)."""
        c = coords.reshape(-1, 2).to(device="cuda", dtype=torch.float32).contiguous()
        if not assume_valid:
            assert c.min() >= 0 and c.max() <= 1, "coordinates should be in [0, 1]"
        out = torch.empty((c.shape[0], self.C), device="cuda", dtype=torch.float32)
        check(lib().dvt_fit_query(self._h, ptr(c), c.shape[0], ptr(out), cur_stream()), "dvt_fit_query")
        return out.reshape(*coords.shape[:-1], self.C)

    def residual(self, raw: torch.Tensor) -> torch.Tensor:
        r = raw.reshape(-1, self.C).to(device="cuda", dtype=torch.float32).contiguous()
        out = torch.empty_like(r)
        check(lib().dvt_fit_residual(self._h, ptr(r), r.shape[0], ptr(out), cur_stream()), "dvt_fit_residual")
        return out.reshape(raw.shape)

    def sweep_once(self, ctas: int = 0):
        """Measurement hook (bench.py, ncu): one dense Adam sweep of the table on the current stream."""
        check(lib().dvt_fit_sweep_once(self._h, ctas, cur_stream()), "dvt_fit_sweep_once")

    def fit(self, denoiser, neural_field, bank_feats, bank_coords, idx_stream, *, graph_steps: int = 10, **hyper):
        """Whole per-image fit; returns the per-step loss table [num_iters, 5]."""
        self.load_modules(denoiser, neural_field)
        self.begin(bank_feats, bank_coords, idx_stream, **hyper)
        self.run(graph_steps=graph_steps)
        return self


def make_patch_coordinates(height, width, start=-1, end=1):
    """reference main_img_denoising.py:21-25"""
    patch_y, patch_x = torch.linspace(start, end, height), torch.linspace(start, end, width)
    patch_y, patch_x = torch.meshgrid(patch_y, patch_x, indexing="ij")
    return torch.stack([patch_x, patch_y], dim=-1)
