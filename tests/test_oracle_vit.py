"""Pins oracle/vit.py against an independent implementation (transformers Dinov2Model; fixtures made by
tests/golden/make_vit_golden.py) and checks the wrapper-level contracts the reference relies on."""
import os

import numpy as np
import pytest
import torch

from oracle import vit as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    e, d, h, p, img, hid, swiglu, reg = [int(v) for v in z["meta"]]
    cfg = O.ViTConfig(embed_dim=e, depth=d, num_heads=h, patch_size=p, native_img=img, mlp_hidden=hid,
                      swiglu=bool(swiglu), num_reg_tokens=reg, no_embed_class=reg > 0)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    return cfg, sd, torch.from_numpy(z["x"]), torch.from_numpy(z["hf_last_hidden_state"])


@pytest.mark.parametrize("name", ["vit_hf_dinov2.npz", "vit_hf_dinov2_reg4.npz", "vit_hf_dinov2_swiglu.npz"])
def test_oracle_matches_hf_dinov2(name):
    cfg, sd, x, ref = _load(name)
    feat, prefix = O.forward_intermediates(sd, cfg, x, [cfg.depth - 1], norm=True, reshape=False,
                                           return_prefix_tokens=True)[0]
    got = torch.cat([prefix, feat], dim=1)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), (got - ref).abs().max()


def test_reshape_and_stride_contract():
    cfg = O.ViTConfig(embed_dim=64, depth=2, num_heads=2, patch_size=14, native_img=56, mlp_hidden=256)
    sd = O.random_state_dict(cfg, seed=3)
    x = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(0))
    nchw = O.forward_intermediates(sd, cfg, x, [1])[0]
    nlc = O.forward_intermediates(sd, cfg, x, [1], reshape=False)[0]
    assert nchw.shape == (2, 64, 4, 4)
    assert torch.equal(nchw.permute(0, 2, 3, 1).reshape(2, 16, 64), nlc)
    # overlapping patches (stride 7): (56-14)//7+1 = 7 -> pos-embed resampled 4x4 -> 7x7
    s7 = O.forward_intermediates(sd, cfg, x, [1], stride=7)[0]
    assert s7.shape == (2, 64, 7, 7)
    # intermediate layer == running fewer blocks
    a = O.forward_intermediates(sd, cfg, x, [0])[0]
    b = O.forward_intermediates(sd, cfg, x, [0, 1])[0]
    assert torch.equal(a, b)


def test_oracle_block_matches_reference_vitdet_code():
    """oracle/vit.py attention / block / pos-embed add against outputs of the reference's OWN code for them
    (evaluation/vitdet/vision_transformer.py:69-138, executed by tests/golden/make_vit_block_golden.py on torch layers with
    timm's attribute names): written-out softmax branch, SDPA branch, LayerScale, residual order."""
    z = np.load(os.path.join(GOLD, "vit_ref_block.npz"))
    dim, heads, hidden, B, H, W = [int(v) for v in z["meta"]]
    cfg = O.ViTConfig(embed_dim=dim, depth=1, num_heads=heads, patch_size=14, native_img=14 * H, mlp_hidden=hidden)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    x = torch.from_numpy(z["x"]).reshape(B, H * W, dim)
    got = O.block(x, sd, 0, cfg).reshape(B, H, W, dim)
    for key in ("y_block", "y_block_sdpa"):
        ref = torch.from_numpy(z[key])
        assert torch.allclose(got, ref, atol=5e-6, rtol=1e-5), (key, (got - ref).abs().max())
    p = "blocks.0."
    xn = torch.nn.functional.layer_norm(x, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
    att = O.attention(xn, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], sd[p + "attn.proj.weight"],
                      sd[p + "attn.proj.bias"], heads).reshape(B, H, W, dim)
    assert torch.allclose(att, torch.from_numpy(z["y_attn"]), atol=5e-6, rtol=1e-5)
    # pos-embed of the patch tokens on the stored grid (vision_transformer.py:120-138): prefix position dropped, added as is
    pos = O.resample_abs_pos_embed(torch.from_numpy(z["pos_embed"]), (H, W), (H, W), 1)
    assert torch.equal(x + pos[:, 1:], torch.from_numpy(z["y_pos"]).reshape(B, H * W, dim))
