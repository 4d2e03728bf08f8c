"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the multiresolution hash-grid encoding that DVT instantiates through tiny-cuda-nn:
`tcnn.Encoding(n_input_dims=2, {"otype": "HashGrid", n_levels, n_features_per_level=8, log2_hashmap_size=20,
base_resolution=16, per_level_scale, "interpolation": "linear"}, dtype=float32)` (reference
dvt/models/neural_feature_field.py:25-39, forward at :46-49).

tiny-cuda-nn is a third-party dependency that is NOT under /root/reference and is unpinned there (README.md:63
installs master); its published algorithm (include/tiny-cuda-nn/encodings/grid.h: grid_scale, grid_resolution,
grid_index, coherent-prime hash, pos_fract, kernel_grid) is restated here:
    scale_l = exp2f(l * log2f(per_level_scale)) * base_resolution - 1          (fp32)
    res_l   = ceilf(scale_l) + 1
    size_l  = min(round_up(res_l^2, 8), 2^log2_hashmap_size)                    (entries of F floats)
    pos     = fmaf(scale_l, x, 0.5); cell = floor(pos); w = pos - cell          (per input dim)
    index   = cx + cy * res_l           if res_l^2 <= size_l (dense)
              (cx * 1) xor (cy * 2654435761)  (uint32)  otherwise (hashed)      then  % size_l
    out[:, l*F:(l+1)*F] = sum over the 4 corners of weight * table[offset_l + index]

Known residual difference to a real tcnn run: tcnn evaluates `grid_scale` with the DEVICE exp2f inside its kernels (CUDA
exp2f: up to 2 ulp) and with the host libm for the offset table; this restatement (and the CUDA path, which reads the scale
from the same host-computed table) uses the correctly rounded host value for both.  A 1-2 ulp difference of scale_l moves
`pos` by <= 2^-22 relative: interpolation weights change by ~1e-6, and a coordinate lands in the neighbouring cell only
when it sits within that distance of a cell boundary.

PARITY UNPINNED: tiny-cuda-nn cannot be imported here and the reference holds no golden vectors for it
(SURVEY.md section 8c).  What IS pinned is the CUDA path against this restatement (bit-exact indices and
interpolation weights, tests/test_fit_gpu.py) and the level table printed in SURVEY.md section 8a-5.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

PRIME_Y = np.uint32(2654435761)


@dataclass
class GridMeta:
    n_levels: int
    n_features: int
    per_level_scale: float
    scale: np.ndarray      # float32 [L]
    res: np.ndarray        # uint32 [L]
    size: np.ndarray       # uint32 [L] entries per level
    offset: np.ndarray     # uint32 [L+1] entry offsets
    hashed: np.ndarray     # bool [L]

    @property
    def n_entries(self) -> int:
        return int(self.offset[-1])

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


def grid_meta(n_levels: int = 16, base_resolution: int = 16, max_resolution: int = 1024, n_features: int = 8,
              log2_hashmap_size: int = 20) -> GridMeta:
    # reference neural_feature_field.py:34-36 computes per_level_scale in float64; tcnn stores it as float
    pls64 = float(np.exp((np.log(max_resolution) - np.log(base_resolution)) / (n_levels - 1))) if n_levels > 1 else 1.0
    pls = np.float32(pls64)
    log2_pls = np.log2(pls).astype(np.float32)
    scale = np.zeros(n_levels, np.float32)
    res = np.zeros(n_levels, np.uint32)
    size = np.zeros(n_levels, np.uint32)
    hashed = np.zeros(n_levels, bool)
    offset = np.zeros(n_levels + 1, np.uint32)
    cap = np.uint64(1) << np.uint64(log2_hashmap_size)
    for l in range(n_levels):
        s = np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(base_resolution) - np.float32(1.0)
        scale[l] = s
        r = np.uint32(np.ceil(s)) + np.uint32(1)
        res[l] = r
        dense = np.uint64(r) * np.uint64(r)
        params = (dense + np.uint64(7)) // np.uint64(8) * np.uint64(8)
        size[l] = min(params, cap)
        hashed[l] = dense > np.uint64(size[l])
        offset[l + 1] = offset[l] + size[l]
    return GridMeta(n_levels, n_features, pls64, scale, res, size, offset, hashed)


def corner_indices_weights(coords: torch.Tensor, meta: GridMeta, level: int):
    """coords fp32 [N,2] in [0,1] -> (idx int64 [N,4] entry index inside the level, w fp32 [N,4]).
    Corner order: (dx,dy) = (0,0),(1,0),(0,1),(1,1)  (tcnn: bit `dim` of the corner id selects +1 on that dim)."""
    s = float(meta.scale[level])
    pos = (coords.double() * s + 0.5).float()        # == fmaf(scale, x, 0.5) in fp32 (single rounding)
    cell = torch.floor(pos)
    w = pos - cell
    c = cell.to(torch.int64)
    res, size = int(meta.res[level]), int(meta.size[level])
    idx, wt = [], []
    for dy in (0, 1):
        for dx in (0, 1):
            x, y = c[:, 0] + dx, c[:, 1] + dy
            if meta.hashed[level]:
                hx = x & 0xFFFFFFFF
                hy = (y * int(PRIME_Y)) & 0xFFFFFFFF
                i = (hx ^ hy) % size
            else:
                i = (x + y * res) % size
            idx.append(i)
            wx = w[:, 0] if dx else 1.0 - w[:, 0]
            wy = w[:, 1] if dy else 1.0 - w[:, 1]
            wt.append(wx * wy)
    # reorder to tcnn corner id order: id bit0 = dx, bit1 = dy -> (0,0),(1,0),(0,1),(1,1): already that order
    return torch.stack(idx, 1), torch.stack(wt, 1)


def encode(table: torch.Tensor, coords: torch.Tensor, meta: GridMeta) -> torch.Tensor:
    """table fp32 [n_entries, F] (or flat) -> [N, L*F].  Differentiable w.r.t. `table` (dense gradient, like tcnn)."""
    F_ = meta.n_features
    t = table.reshape(-1, F_)
    outs: List[torch.Tensor] = []
    for l in range(meta.n_levels):
        idx, w = corner_indices_weights(coords, meta, l)
        rows = t[int(meta.offset[l]) + idx]            # [N,4,F]
        outs.append((rows * w.unsqueeze(-1)).sum(1))
    return torch.cat(outs, dim=1)
