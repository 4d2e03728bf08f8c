"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the stage-2 `Denoiser` FORWARD (inference):
  * `Denoiser.forward`      dvt/models/online_denoiser.py:62-104
  * timm `Block` (pre-LN, qkv_bias, no LayerScale: `init_values=None`, GELU Mlp, LayerNorm eps 1e-6), constructed at
    online_denoiser.py:25-52; its maths is the block of oracle/vit.py (pinned against `transformers.Dinov2Model`,
    tests/golden/vit_hf_*.npz) with the LayerScale factors absent
  * `timm.layers.resample_abs_pos_embed(pos_embed, (h, w), num_prefix_tokens=0)` = oracle/vit.py::resample_abs_pos_embed

PINNING: the reference class cannot be imported here (it needs timm, which is not installed): parity unpinned against
timm itself; pinned transitively through the block of oracle/vit.py."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import vit as OV


def random_state_dict(feat_dim: int, hw: Tuple[int, int], num_blocks: int = 1, enable_pe: bool = True, seed: int = 0
                      ) -> Dict[str, torch.Tensor]:
    """State dict with the reference's key names (`denoiser.<timm Block names>` / `denoiser.<i>.…`, `pos_embed`)."""
    g = torch.Generator().manual_seed(seed)
    C, Hd = feat_dim, 4 * feat_dim
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sd: Dict[str, torch.Tensor] = {}
    for i in range(num_blocks):
        p = "denoiser." if num_blocks <= 1 else f"denoiser.{i}."
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = 1 + 0.1 * rn(C), 0.1 * rn(C)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = rn(3 * C, C) * C ** -0.5, 0.1 * rn(3 * C)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = rn(C, C) * C ** -0.5, 0.1 * rn(C)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = 1 + 0.1 * rn(C), 0.1 * rn(C)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = rn(Hd, C) * C ** -0.5, 0.1 * rn(Hd)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = rn(C, Hd) * Hd ** -0.5, 0.1 * rn(C)
    if enable_pe:
        sd["pos_embed"] = rn(1, hw[0] * hw[1], C) * 0.02
    return sd


def forward(sd: Dict[str, torch.Tensor], x_bhwc: torch.Tensor, noise_map_hw: Tuple[int, int], num_blocks: int = 1
            ) -> torch.Tensor:
    """online_denoiser.py:86-92 for a feature map x [b, h, w, c] (the `vit is None` branch; with a ViT the caller feeds
    its NHWC output)."""
    b, h, w, c = x_bhwc.shape
    cfg = OV.ViTConfig(embed_dim=c, depth=num_blocks, num_heads=c // 64, mlp_hidden=4 * c, layerscale=False)
    x = x_bhwc.reshape(b, h * w, c).float()
    pos: Optional[torch.Tensor] = sd.get("pos_embed")
    if pos is not None:
        x = x + OV.resample_abs_pos_embed(pos.float(), (h, w), noise_map_hw, 0)
    for i in range(num_blocks):
        p = "denoiser." if num_blocks <= 1 else f"denoiser.{i}."
        blk = {k.replace(p, f"blocks.{i}."): v for k, v in sd.items() if k.startswith(p)}
        x = OV.block(x, blk, i, cfg)
    return x.reshape(b, h, w, c)
