"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the frozen-ViT forward that DVT reaches through
`PretrainedViTWrapper.get_intermediate_layers` (reference: dvt/models/vit_wrapper.py:122-143), i.e. timm 1.0.7
`VisionTransformer.forward_intermediates` with the wrapper's stride override (vit_wrapper.py:78-91).

timm is a third-party dependency that is NOT vendored under /root/reference (requirements.txt:3 pins
timm==1.0.7) and is not installed in this image, so its published algorithm is restated here; the in-repo
restatements the reference itself carries are followed where they exist:
  * attention            -> evaluation/vitdet/vision_transformer.py:73-91
  * block order, ls1/ls2 -> evaluation/vitdet/vision_transformer.py:98-117
  * pos-embed handling   -> evaluation/vitdet/vision_transformer.py:120-138

PINNING: the reference has no tests or golden vectors for this path.  This restatement is pinned
(i) against the reference's OWN code for attention, block order / LayerScale / residuals and the pos-embed add --
evaluation/vitdet/vision_transformer.py:69-138 executed on torch layers (tests/golden/make_vit_block_golden.py ->
tests/golden/vit_ref_block.npz, both attention branches) -- and (ii) end to end against an independent implementation of the
same architecture, `transformers.Dinov2Model` (tests/golden/make_vit_golden.py -> tests/golden/vit_hf_*.npz); both are
checked by tests/test_oracle_vit.py.  The pos-embed *resampling* branch (grid != native) has no second implementation
available offline: "parity unpinned" for that branch only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class ViTConfig:
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    patch_size: int = 14
    native_img: int = 518          # training resolution that fixes the stored pos-embed grid
    mlp_hidden: int = 3072
    swiglu: bool = False           # timm SwiGLUPacked + SiLU (ViT-g/14)
    layerscale: bool = True        # ls1 / ls2 present (DINOv2 init_values=1e-5)
    num_reg_tokens: int = 0
    no_embed_class: bool = False   # True: pos_embed covers patches only (reg4 models)
    ln_eps: float = 1e-6

    @property
    def native_grid(self) -> int:
        return self.native_img // self.patch_size

    @property
    def num_prefix(self) -> int:
        return 1 + self.num_reg_tokens


# timm 1.0.7 model definitions for the DINOv2 family used by the headline configs (MODEL_LIST,
# dvt/models/vit_wrapper.py:15-56).  Other families keep the same block structure with different sizes.
CONFIGS: Dict[str, ViTConfig] = {
    "vit_small_patch14_dinov2.lvd142m": ViTConfig(384, 12, 6, 14, 518, 1536),
    "vit_base_patch14_dinov2.lvd142m": ViTConfig(768, 12, 12, 14, 518, 3072),
    "vit_large_patch14_dinov2.lvd142m": ViTConfig(1024, 24, 16, 14, 518, 4096),
    "vit_giant_patch14_dinov2.lvd142m": ViTConfig(1536, 40, 24, 14, 518, 8192, swiglu=True),
    "vit_small_patch14_reg4_dinov2.lvd142m": ViTConfig(384, 12, 6, 14, 518, 1536, num_reg_tokens=4, no_embed_class=True),
    "vit_base_patch14_reg4_dinov2.lvd142m": ViTConfig(768, 12, 12, 14, 518, 3072, num_reg_tokens=4, no_embed_class=True),
    "vit_large_patch14_reg4_dinov2.lvd142m": ViTConfig(1024, 24, 16, 14, 518, 4096, num_reg_tokens=4, no_embed_class=True),
    "vit_giant_patch14_reg4_dinov2.lvd142m": ViTConfig(1536, 40, 24, 14, 518, 8192, swiglu=True, num_reg_tokens=4,
                                                       no_embed_class=True),
}


def random_state_dict(cfg: ViTConfig, seed: int = 0, layerscale_range: Tuple[float, float] = (0.5, 1.5)) -> Dict[str, torch.Tensor]:
    """timm-named random weights.  LayerScale is drawn from U(0.5, 1.5) instead of DINOv2's 1e-5 so that parity
    tests actually exercise the attention / MLP branches ('hard parts')."""
    g = torch.Generator().manual_seed(seed)
    C, P = cfg.embed_dim, cfg.patch_size

    def tn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd = {
        "cls_token": tn(1, 1, C),
        "pos_embed": tn(1, cfg.native_grid ** 2 + (0 if cfg.no_embed_class else 1), C),
        "patch_embed.proj.weight": tn(C, 3, P, P, std=1.0 / math.sqrt(3 * P * P)),
        "patch_embed.proj.bias": tn(C),
        "norm.weight": 1.0 + tn(C, std=0.1),
        "norm.bias": tn(C, std=0.1),
    }
    if cfg.num_reg_tokens:
        sd["reg_token"] = tn(1, cfg.num_reg_tokens, C)
    lo, hi = layerscale_range
    fc2_in = cfg.mlp_hidden // 2 if cfg.swiglu else cfg.mlp_hidden
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + tn(C, std=0.1)
        sd[p + "norm1.bias"] = tn(C, std=0.1)
        sd[p + "attn.qkv.weight"] = tn(3 * C, C, std=1.0 / math.sqrt(C))
        sd[p + "attn.qkv.bias"] = tn(3 * C)
        sd[p + "attn.proj.weight"] = tn(C, C, std=0.5 / math.sqrt(C))
        sd[p + "attn.proj.bias"] = tn(C)
        sd[p + "norm2.weight"] = 1.0 + tn(C, std=0.1)
        sd[p + "norm2.bias"] = tn(C, std=0.1)
        sd[p + "mlp.fc1.weight"] = tn(cfg.mlp_hidden, C, std=1.0 / math.sqrt(C))
        sd[p + "mlp.fc1.bias"] = tn(cfg.mlp_hidden)
        sd[p + "mlp.fc2.weight"] = tn(C, fc2_in, std=0.5 / math.sqrt(fc2_in))
        sd[p + "mlp.fc2.bias"] = tn(C)
        if cfg.layerscale:
            sd[p + "ls1.gamma"] = lo + (hi - lo) * torch.rand(C, generator=g)
            sd[p + "ls2.gamma"] = lo + (hi - lo) * torch.rand(C, generator=g)
    return sd


def feat_size(cfg: ViTConfig, height: int, width: int, stride: int) -> Tuple[int, int]:
    """dvt/models/vit_wrapper.py:81-87 (the patched dynamic_feat_size)."""
    return (height - cfg.patch_size) // stride + 1, (width - cfg.patch_size) // stride + 1


def resample_abs_pos_embed(pos: torch.Tensor, new_hw: Tuple[int, int], old_hw: Tuple[int, int], num_prefix: int) -> torch.Tensor:
    """timm.layers.resample_abs_pos_embed: bicubic + antialias in fp32 on the patch part, prefix passed through."""
    if tuple(new_hw) == tuple(old_hw):
        return pos
    prefix, grid = pos[:, :num_prefix], pos[:, num_prefix:]
    C = grid.shape[-1]
    grid = grid.reshape(1, old_hw[0], old_hw[1], C).permute(0, 3, 1, 2).float()
    grid = F.interpolate(grid, size=new_hw, mode="bicubic", antialias=True)
    grid = grid.permute(0, 2, 3, 1).reshape(1, -1, C)
    return torch.cat([prefix, grid], dim=1) if num_prefix else grid


def attention(x: torch.Tensor, qkv_w, qkv_b, proj_w, proj_b, num_heads: int) -> torch.Tensor:
    """evaluation/vitdet/vision_transformer.py:73-91 (non-fused branch written out)."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, qkv_w, qkv_b).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, proj_w, proj_b)


def mlp(x: torch.Tensor, sd, p: str, cfg: ViTConfig) -> torch.Tensor:
    h = F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    if cfg.swiglu:
        x1, x2 = h.chunk(2, dim=-1)      # timm GluMlp(gate_last=False): act(x1) * x2
        h = F.silu(x1) * x2
    else:
        h = F.gelu(h)                    # erf GELU
    return F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def block(x: torch.Tensor, sd, i: int, cfg: ViTConfig) -> torch.Tensor:
    """evaluation/vitdet/vision_transformer.py:98-117."""
    p = f"blocks.{i}."
    C = cfg.embed_dim
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
    y = attention(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], sd[p + "attn.proj.weight"],
                  sd[p + "attn.proj.bias"], cfg.num_heads)
    if cfg.layerscale:
        y = y * sd[p + "ls1.gamma"]
    x = x + y
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
    y = mlp(y, sd, p, cfg)
    if cfg.layerscale:
        y = y * sd[p + "ls2.gamma"]
    return x + y


def embed(sd, cfg: ViTConfig, x: torch.Tensor, stride: int) -> torch.Tensor:
    """patch_embed (conv with the overridden stride, NHWC) + _pos_embed of timm 1.0.7 with dynamic_img_size."""
    B, _, H, W = x.shape
    h, w = feat_size(cfg, H, W, stride)
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=stride)
    assert t.shape[-2:] == (h, w)
    t = t.permute(0, 2, 3, 1).reshape(B, h * w, cfg.embed_dim)
    npre_pos = 0 if cfg.no_embed_class else 1
    pos = resample_abs_pos_embed(sd["pos_embed"], (h, w), (cfg.native_grid, cfg.native_grid), npre_pos)
    to_cat = [sd["cls_token"].expand(B, -1, -1)]
    if cfg.num_reg_tokens:
        to_cat.append(sd["reg_token"].expand(B, -1, -1))
    if cfg.no_embed_class:
        t = t + pos
        t = torch.cat(to_cat + [t], dim=1)
    else:
        t = torch.cat(to_cat + [t], dim=1)
        t = t + pos
    return t


@torch.no_grad()
def forward_intermediates(sd: Dict[str, torch.Tensor], cfg: ViTConfig, x: torch.Tensor, indices: Sequence[int],
                          stride: int | None = None, norm: bool = True, reshape: bool = True,
                          return_prefix_tokens: bool = False) -> List[torch.Tensor]:
    """`PretrainedViTWrapper.get_intermediate_layers(x, n=indices, reshape, return_prefix_tokens, norm)`.

    All blocks are NOT needed past max(indices); the reference runs them anyway (no stop_early,
    vit_wrapper.py:136-143) without effect on the returned tensors.
    """
    stride = cfg.patch_size if stride is None else stride
    B, _, H, W = x.shape
    h, w = feat_size(cfg, H, W, stride)
    t = embed(sd, cfg, x.float(), stride)
    outs = []
    for i in range(max(indices) + 1):
        t = block(t, sd, i, cfg)
        if i in indices:
            y = F.layer_norm(t, (cfg.embed_dim,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps) if norm else t
            outs.append(y)
    res = []
    for y in outs:
        prefix, feat = y[:, :cfg.num_prefix], y[:, cfg.num_prefix:]
        if reshape:
            feat = feat.reshape(B, h, w, -1).permute(0, 3, 1, 2).contiguous()
        res.append((feat, prefix) if return_prefix_tokens else feat)
    return res
