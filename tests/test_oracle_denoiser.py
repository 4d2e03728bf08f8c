"""CPU: the stage-2 Denoiser forward oracle against an independent torch construction of the same block
(nn.MultiheadAttention + nn.LayerNorm + nn.Linear with the weights copied in), and the state-dict contract of the
drop-in module."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def test_oracle_block_matches_torch_modules():
    from oracle import denoiser as OD
    C, hw = 128, (5, 6)
    sd = OD.random_state_dict(C, hw, num_blocks=1, seed=3)
    x = torch.randn(2, hw[0], hw[1], C, generator=torch.Generator().manual_seed(1))
    got = OD.forward(sd, x, hw)
    # independent construction: torch's own fused attention module
    mha = nn.MultiheadAttention(C, C // 64, bias=True, batch_first=True)
    with torch.no_grad():
        mha.in_proj_weight.copy_(sd["denoiser.attn.qkv.weight"]); mha.in_proj_bias.copy_(sd["denoiser.attn.qkv.bias"])
        mha.out_proj.weight.copy_(sd["denoiser.attn.proj.weight"]); mha.out_proj.bias.copy_(sd["denoiser.attn.proj.bias"])
        t = x.reshape(2, -1, C) + sd["pos_embed"]
        y = F.layer_norm(t, (C,), sd["denoiser.norm1.weight"], sd["denoiser.norm1.bias"], 1e-6)
        t = t + mha(y, y, y, need_weights=False)[0]
        y = F.layer_norm(t, (C,), sd["denoiser.norm2.weight"], sd["denoiser.norm2.bias"], 1e-6)
        y = F.linear(F.gelu(F.linear(y, sd["denoiser.mlp.fc1.weight"], sd["denoiser.mlp.fc1.bias"])),
                     sd["denoiser.mlp.fc2.weight"], sd["denoiser.mlp.fc2.bias"])
        ref = (t + y).reshape(2, hw[0], hw[1], C)
    assert (got - ref).abs().max().item() < 2e-5


def test_denoiser_state_dict_contract():
    """Key names a reference checkpoint has (main_denoiser.py:248-264 saves `denoiser.state_dict()`): timm Block names
    under `denoiser.`, `pos_embed`; `denoiser.<i>.` for num_blocks > 1."""
    import dvt.models as DVT
    from oracle import denoiser as OD
    for nb in (1, 2):
        m = DVT.Denoiser(5, 6, 128, vit=None, enable_pe=True, num_blocks=nb)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in OD.random_state_dict(128, (5, 6), nb).items()}
        assert ours == ref
        m.load_state_dict(OD.random_state_dict(128, (5, 6), nb), strict=True)
    assert DVT.Denoiser(5, 6, 128, enable_pe=False).pos_embed is None


def test_center_padding_matches_reference_rule():
    """`dvt.models.CenterPadding` against the rule of evaluation/eval_utils/misc.py:19-35 (pad each spatial size up to the
    next multiple, left = pad // 2, right = the rest), and against the reference class itself where it can be loaded."""
    import math
    import dvt.models as DVT
    pad = DVT.CenterPadding(14)
    for H, W in [(480, 850), (490, 854), (1, 15), (27, 28), (518, 518)]:
        x = torch.arange(2 * 3 * H * W, dtype=torch.float32).reshape(2, 3, H, W)
        y = pad(x)
        nh, nw = math.ceil(H / 14) * 14, math.ceil(W / 14) * 14
        t, l = (nh - H) // 2, (nw - W) // 2
        assert y.shape == (2, 3, nh, nw)
        assert torch.equal(y[:, :, t:t + H, l:l + W], x)
        assert float(y.sum()) == float(x.sum())          # everything else is zero
    assert pad(torch.ones(1, 2, 5, 6, 7)).shape == (1, 2, 14, 14, 14)   # any number of trailing dimensions
