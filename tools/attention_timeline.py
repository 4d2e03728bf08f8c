#!/usr/bin/env python
"""Per-key-tile milestones inside the attention kernel (CTA (0,0,0); clock64 of its SM): where does a tile's time go?
  DVT_ATTN_DEBUG_TS=1 python tools/attention_timeline.py [--batch 16]"""
import argparse
import os
import sys

os.environ["DVT_ATTN_DEBUG_TS"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
from dvt import ops  # noqa: E402
from dvt._lib import check, lib, ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
B, N, H = a.batch, 1370, 12
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3 * H * 64, device="cuda", generator=g).bfloat16()
ts = torch.zeros(16 + 8 * 16, dtype=torch.int64, device="cuda")
for _ in range(20):
    ops.attention(qkv, H)
torch.cuda.synchronize()
check(lib().dvt_debug_set_timestamp_buffer(ptr(ts)))
for _ in range(3):
    ops.attention(qkv, H)
torch.cuda.synchronize()
check(lib().dvt_debug_set_timestamp_buffer(None))
t = ts.cpu().tolist()
names = ["wait S", "S ready", "S in regs", "P computed", "PV(j-1) done", "P stored", "QK(j+1) issued", "PV(j) issued"]
t0 = t[16]
print("tile " + " ".join(f"{n:>15s}" for n in names) + "   (clk from the first stamp)")
for j in range(11):
    row = t[16 + 8 * j: 24 + 8 * j]
    print(f"{j:4d} " + " ".join(f"{(v - t0) if v else -1:15d}" for v in row))
