"""Generates tests/golden/vit_hf_*.npz: outputs of an INDEPENDENT implementation of the DINOv2 ViT
(`transformers.Dinov2Model` / `Dinov2WithRegistersModel`, random-init, no checkpoint) used to pin oracle/vit.py.

Run in the build container:  python tests/golden/make_vit_golden.py
The .npz holds the input, the timm-named weights (fp16-free, float32) and HF's last_hidden_state.
Sizes are kept tiny (embed 64, 1 head of 64, depth 2, 56x56 image, patch 14 -> 4x4 grid) so the fixture is a few hundred KB.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def hf_to_timm(hf_sd, depth, with_reg):
    sd = {
        "cls_token": hf_sd["embeddings.cls_token"],
        "pos_embed": hf_sd["embeddings.position_embeddings"],
        "patch_embed.proj.weight": hf_sd["embeddings.patch_embeddings.projection.weight"],
        "patch_embed.proj.bias": hf_sd["embeddings.patch_embeddings.projection.bias"],
        "norm.weight": hf_sd["layernorm.weight"],
        "norm.bias": hf_sd["layernorm.bias"],
    }
    if with_reg:
        # timm's reg4 checkpoints are stored with no_embed_class=True: the cls position is folded into cls_token
        # and pos_embed covers the patches only (timm vision_transformer.py checkpoint_filter_fn for DINOv2-reg).
        sd["reg_token"] = hf_sd["embeddings.register_tokens"]
        sd["cls_token"] = sd["cls_token"] + sd["pos_embed"][:, :1]
        sd["pos_embed"] = sd["pos_embed"][:, 1:]
    for i in range(depth):
        h = f"encoder.layer.{i}."
        t = f"blocks.{i}."
        a = h + "attention.attention."
        sd[t + "attn.qkv.weight"] = torch.cat([hf_sd[a + "query.weight"], hf_sd[a + "key.weight"], hf_sd[a + "value.weight"]])
        sd[t + "attn.qkv.bias"] = torch.cat([hf_sd[a + "query.bias"], hf_sd[a + "key.bias"], hf_sd[a + "value.bias"]])
        sd[t + "attn.proj.weight"] = hf_sd[h + "attention.output.dense.weight"]
        sd[t + "attn.proj.bias"] = hf_sd[h + "attention.output.dense.bias"]
        for n in ("norm1", "norm2"):
            sd[t + n + ".weight"] = hf_sd[h + n + ".weight"]
            sd[t + n + ".bias"] = hf_sd[h + n + ".bias"]
        sd[t + "ls1.gamma"] = hf_sd[h + "layer_scale1.lambda1"]
        sd[t + "ls2.gamma"] = hf_sd[h + "layer_scale2.lambda1"]
        if h + "mlp.fc1.weight" in hf_sd:
            sd[t + "mlp.fc1.weight"] = hf_sd[h + "mlp.fc1.weight"]
            sd[t + "mlp.fc1.bias"] = hf_sd[h + "mlp.fc1.bias"]
            sd[t + "mlp.fc2.weight"] = hf_sd[h + "mlp.fc2.weight"]
            sd[t + "mlp.fc2.bias"] = hf_sd[h + "mlp.fc2.bias"]
        else:  # SwiGLU: HF weights_in / weights_out == timm SwiGLUPacked fc1 / fc2
            sd[t + "mlp.fc1.weight"] = hf_sd[h + "mlp.weights_in.weight"]
            sd[t + "mlp.fc1.bias"] = hf_sd[h + "mlp.weights_in.bias"]
            sd[t + "mlp.fc2.weight"] = hf_sd[h + "mlp.weights_out.weight"]
            sd[t + "mlp.fc2.bias"] = hf_sd[h + "mlp.weights_out.bias"]
    return {k: v.detach().float().contiguous() for k, v in sd.items()}


def make(name, with_reg=False, swiglu=False, seed=0):
    import transformers
    torch.manual_seed(seed)
    kw = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=1, mlp_ratio=4, image_size=56, patch_size=14,
              layerscale_value=1.0, use_swiglu_ffn=swiglu, hidden_act="gelu", attention_probs_dropout_prob=0.0,
              hidden_dropout_prob=0.0, layer_norm_eps=1e-6)
    if with_reg:
        cfg = transformers.Dinov2WithRegistersConfig(num_register_tokens=4, **kw)
        model = transformers.Dinov2WithRegistersModel(cfg)
    else:
        cfg = transformers.Dinov2Config(**kw)
        model = transformers.Dinov2Model(cfg)
    model.eval()
    # non-trivial values everywhere (HF init leaves biases at 0 and LayerScale at layerscale_value)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "lambda1" in n:
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif n.endswith("bias") or "token" in n or "position" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
    x = torch.randn(2, 3, 56, 56, generator=g)
    with torch.no_grad():
        out = model(pixel_values=x).last_hidden_state  # [B, prefix + 16, C], final LayerNorm applied
    hf_sd = {k: v for k, v in model.state_dict().items()}
    sd = hf_to_timm(hf_sd, 2, with_reg)
    mlp_hidden = sd["blocks.0.mlp.fc1.weight"].shape[0]
    arrays = {"x": x.numpy(), "hf_last_hidden_state": out.numpy(),
              "meta": np.array([64, 2, 1, 14, 56, mlp_hidden, int(swiglu), 4 if with_reg else 0], dtype=np.int64)}
    for k, v in sd.items():
        arrays["w:" + k] = v.numpy()
    path = os.path.join(HERE, f"vit_hf_{name}.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "transformers", transformers.__version__)


if __name__ == "__main__":
    make("dinov2")
    make("dinov2_reg4", with_reg=True)
    make("dinov2_swiglu", swiglu=True)
