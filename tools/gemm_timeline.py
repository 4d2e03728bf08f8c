#!/usr/bin/env python
"""Where does the time of a one-tile-per-CTA 3xTF32 GEMM go?  %globaltimer milestones of CTA 0 (ns from kernel entry)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
from dvt import ops  # noqa: E402
from dvt._lib import check, lib, ptr  # noqa: E402

ts = torch.zeros(16, dtype=torch.int64, device="cuda")
check(lib().dvt_debug_set_timestamp_buffer(ptr(ts)))
g = torch.Generator(device="cuda").manual_seed(0)
names = ["entry", "setup", "operands", "mma issued", "epi start", "epi end", "exit"]
for (M, N, K, amn, bmn, label) in [(2048, 384, 128, False, False, "G1 fwd K=128"), (2048, 768, 384, False, False, "G2 fwd K=384"),
                                   (2048, 384, 768, False, True, "dgrad K=768 (B MN)"), (2048, 128, 384, False, True, "dgrad K=384 (B MN)")]:
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(K, N, device="cuda", generator=g) if bmn else torch.randn(N, K, device="cuda", generator=g)
    for _ in range(3):
        ops.gemm_f32x3(a, b, M, N, K, a_mn=amn, b_mn=bmn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ap, bp = ops.split_tf32(a), ops.split_tf32(b)
    out = torch.empty(M, N, device="cuda")
    for _ in range(200):  # keep the GPU busy so that the clocks are up when the measured launch runs
        check(lib().dvt_gemm_f32x3(ptr(ap), a.shape[1], a.numel(), int(amn), ptr(bp), b.shape[1], b.numel(), int(bmn), M, N, K,
                                   ptr(out), N, 1, None, None))
    e0.record()
    check(lib().dvt_gemm_f32x3(ptr(ap), a.shape[1], a.numel(), int(amn), ptr(bp), b.shape[1], b.numel(), int(bmn), M, N, K,
                               ptr(out), N, 1, None, None))
    e1.record()
    torch.cuda.synchronize()
    t = ts.cpu().tolist()
    rel = [f"{n}={t[i] - t[0]:6d}" for i, n in enumerate(names)]
    print(f"{label:22s} event {e0.elapsed_time(e1) * 1e3:7.1f} us | ns: " + "  ".join(rel))
    print("    epilogue warp 0 chunk completions (ns from entry):", [t[8 + k] - t[0] for k in range(4)],
          f" SM clock during kernel: {(t[15] - t[14]) / max(t[6] - t[0], 1) * 1e3:.0f} MHz")
check(lib().dvt_debug_set_timestamp_buffer(None))
