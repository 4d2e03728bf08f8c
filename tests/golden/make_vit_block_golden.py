"""Generates tests/golden/vit_ref_block.npz by EXECUTING the reference's own restatement of the timm ViT block:
`attn_forward` (both its SDPA and its written-out softmax branch), the block forward returned by
`get_vit_forward_fn(0)` and `vit_pos_embed` of /root/reference/evaluation/vitdet/vision_transformer.py:69-138 -- the code
the reference monkey-patches onto timm's modules for its own evaluation.  oracle/vit.py cites exactly these lines; this
fixture pins it to them (tests/test_oracle_vit.py::test_oracle_block_matches_reference_vitdet_code).

The module imports mmcv / mmdet / timm at the top for its detector wrapper class only; none of them is installed here, so
the three imported names are stubbed before loading the file (the functions under test do not touch them, except
`resample_abs_pos_embed`, which `vit_pos_embed` calls with the stored grid in this fixture: timm returns the embedding
unchanged in that case, and so does the stub, which refuses any other size).  The layers the functions are applied to are plain
torch.nn modules with timm's attribute names.

Run in the build container (needs /root/reference):  python tests/golden/make_vit_block_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/evaluation/vitdet/vision_transformer.py"


def load_reference_module():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Registry:
        def register_module(self):
            return lambda cls: cls

    def _no_resample(posemb, new_size, num_prefix_tokens=1, **k):
        # timm 1.0.7 returns the embedding unchanged when the grid already has the requested size; the fixture only uses
        # that case (the resampling arithmetic itself is not restated here)
        n = posemb.shape[1] - num_prefix_tokens
        assert n == new_size[0] * new_size[1] and int(n ** 0.5) ** 2 == n, "fixture must use the stored (square) grid"
        return posemb

    stub("mmcv")
    stub("mmcv.runner", BaseModule=nn.Module)
    stub("mmdet")
    stub("mmdet.models")
    stub("mmdet.models.builder", BACKBONES=_Registry())
    stub("timm")
    stub("timm.layers", resample_abs_pos_embed=_no_resample)
    spec = importlib.util.spec_from_file_location("ref_vitdet_vision_transformer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class LayerScale(nn.Module):          # timm LayerScale: x * gamma
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Mlp(nn.Module):                 # timm Mlp(act_layer=nn.GELU): fc1 -> GELU (erf) -> fc2
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):           # attribute names of timm's Attention; forward = the reference's attn_forward
    def __init__(self, dim, heads, ref, fused):
        super().__init__()
        self.num_heads, self.head_dim, self.scale, self.fused_attn = heads, dim // heads, (dim // heads) ** -0.5, fused
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.q_norm = self.k_norm = nn.Identity()
        self.attn_drop = self.proj_drop = nn.Dropout(0.0)
        self._ref = ref

    def forward(self, x):
        return self._ref.attn_forward(self, x)


class Block(nn.Module):               # attribute names of timm's Block; forward = the reference's get_vit_forward_fn(0)
    def __init__(self, dim, heads, hidden, ref, fused):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6), nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads, ref, fused)
        self.ls1, self.ls2 = LayerScale(dim), LayerScale(dim)
        self.drop_path1 = self.drop_path2 = nn.Identity()
        self.mlp = Mlp(dim, hidden)
        self._fwd = ref.get_vit_forward_fn(0)

    def forward(self, x):
        return self._fwd(self, x)


class PosEmbedHolder(nn.Module):      # the attributes vit_pos_embed reads from timm's VisionTransformer
    def __init__(self, dim, h, w, num_prefix):
        super().__init__()
        self.no_embed_class, self.num_prefix_tokens, self.dynamic_img_size = False, num_prefix, True
        self.pos_embed = nn.Parameter(torch.randn(1, num_prefix + h * w, dim) * 0.02)
        self.pos_drop = nn.Identity()


def main():
    ref = load_reference_module()
    torch.manual_seed(0)
    dim, heads, hidden, B, H, W = 128, 2, 256, 2, 5, 5
    out = {"meta": np.array([dim, heads, hidden, B, H, W])}
    x = torch.randn(B, H, W, dim)
    blk = Block(dim, heads, hidden, ref, fused=False)
    with torch.no_grad():
        for p in blk.parameters():       # non-degenerate LayerNorm / LayerScale / biases
            p.copy_(torch.randn_like(p) * (0.5 if p.ndim == 1 else 0.08))
        blk.norm1.weight.add_(1.0)
        blk.norm2.weight.add_(1.0)
        y_block = blk(x)
        y_attn = blk.attn(blk.norm1(x))
        blk.attn.fused_attn = True       # the SDPA branch of the same function
        y_block_sdpa = blk(x)
        pe = PosEmbedHolder(dim, H, W, num_prefix=1)
        y_pos = ref.vit_pos_embed(pe, x)
    out["x"] = x.numpy()
    out["y_block"], out["y_block_sdpa"], out["y_attn"], out["y_pos"] = (t.numpy() for t in (y_block, y_block_sdpa, y_attn, y_pos))
    out["pos_embed"] = pe.pos_embed.detach().numpy()
    for k, v in blk.state_dict().items():
        out["w:blocks.0." + k] = v.numpy()
    path = os.path.join(HERE, "vit_ref_block.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; |block - block_sdpa| max", float((y_block - y_block_sdpa).abs().max()))


if __name__ == "__main__":
    main()
