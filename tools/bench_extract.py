#!/usr/bin/env python
"""HP-1 alone -- BASELINE.json configs[1] and configs[4] plus the long-sequence case of SURVEY.md 8(f-4):
  * extraction only, DINOv2 ViT-B/14, batch 64 synthetic 518 x 518 (config 2), next to the unfused torch restatement
    (SDPA + cuBLAS, fp32 and bf16 autocast: tools/library_bar.py) on the same GPU;
  * ViT-L/14 and ViT-g/14 (SwiGLU) at 518^2, batch 16 (config 5);
  * ViT-B/14 at stride 4 on a 490 x 854 frame: 25 321 tokens per image (make_video_demo.py:21-22).
Per config: views/s, achieved TFLOP/s (SURVEY.md 8(d) FLOPs per view) and the fraction of the measured bf16 peak
(MEASURED_PEAKS.json, sustained figure: these are seconds-long loops).  One JSON line per config on stdout.
CUDA events on the launching stream, 3 warm-up + `--reps` timed forwards, inputs larger than L2."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "denoising-vit_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import bench  # noqa: E402  (peaks())


def vit_flops(embed, depth, heads, mlp_hidden, swiglu, tokens, patches, patch=14):
    """2*M*N*K over patch-embed, QKV, QK^T, PV, out-proj, MLP for one image (SURVEY.md 8(d) accounting)."""
    C = embed
    fc2_in = mlp_hidden // 2 if swiglu else mlp_hidden
    per_block = 2 * tokens * C * 3 * C + 4 * tokens * tokens * C + 2 * tokens * C * C + 2 * tokens * C * mlp_hidden + 2 * tokens * fc2_in * C
    return 2 * patches * 3 * patch * patch * C + depth * per_block


def run(ident, B, H, W, stride, reps, dev):
    import dvt.models as DVT
    from dvt.models import vit_wrapper as VW
    a = VW.ARCHS[ident]
    vit = DVT.PretrainedViTWrapper(ident, stride=stride, allow_random_init=True)
    with torch.no_grad():
        for b in vit.model.blocks:
            if hasattr(b, "ls1") and hasattr(b.ls1, "gamma"):
                b.ls1.gamma.fill_(1.0)
                b.ls2.gamma.fill_(1.0)
    vit = vit.to(dev).eval()
    h, w = (H - 14) // stride + 1, (W - 14) // stride + 1
    x = torch.randn(B, 3, H, W, device=dev).bfloat16()
    out = torch.empty((B, h, w, a["embed"]), device=dev, dtype=torch.float32)
    layer = a["depth"] - 1
    for _ in range(3):
        vit.extract_into(x, layer, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        vit.extract_into(x, layer, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tokens = h * w + 1 + a.get("reg", 0)
    fl = vit_flops(a["embed"], a["depth"], a["heads"], a["mlp"], a["swiglu"], tokens, h * w) * B
    pk = bench.peaks()
    tf = fl / (ms / 1e3) / 1e12
    return {"model": ident, "batch": B, "input": [H, W], "stride": stride, "tokens_per_image": tokens, "ms_per_forward": ms,
            "views_per_s": B / (ms / 1e3), "tflops": tf, "gflop_per_view": fl / B / 1e9, "peak_tflops": pk["bf16_tflops"],
            "frac_of_peak": tf / pk["bf16_tflops"], "peak_source": pk["source"] + " (sustained bf16 cuBLAS)", "dtype": "bf16",
            "finite": bool(torch.isfinite(out).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", type=str, default="")
    ap.add_argument("--no-library", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cases = [("config2_vitb_b64", "vit_base_patch14_dinov2.lvd142m", 64, 518, 518, 14),
             ("config5_vitl_b16", "vit_large_patch14_dinov2.lvd142m", 16, 518, 518, 14),
             ("config5_vitg_b16", "vit_giant_patch14_dinov2.lvd142m", 16, 518, 518, 14),
             ("f4_vitb_stride4_490x854", "vit_base_patch14_dinov2.lvd142m", 1, 490, 854, 4)]
    for name, ident, B, H, W, s in cases:
        if args.only and args.only not in name:
            continue
        r = run(ident, B, H, W, s, args.reps, dev)
        r["case"] = name
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if not args.no_library and (not args.only or "library" in args.only or "config2" in args.only):
        import library_bar
        v = library_bar.vit_forward_bar(dev, batch=64, reps=2)
        print(json.dumps({"case": "config2_library_bar_torch_sdpa_cublas_b64", "s_per_view": v,
                          "views_per_s": {k: 1.0 / t for k, t in v.items()},
                          "tflops": {k: 303.1e9 / t / 1e12 for k, t in v.items()}}), flush=True)


if __name__ == "__main__":
    main()
