#!/usr/bin/env python
"""Do two 3xTF32 GEMMs that fit on the GPU side by side (96 + 48 CTAs <= 148 SMs) slow each other down?  If they do, the
shared resource is the L2 -> SM operand traffic, not the SMs.  Also: the same GEMM beside a full-rate memory stream (a
device-to-device copy), the stand-in for the dense Adam sweep.  CUDA events; 50 repetitions each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))
from dvt import ops  # noqa: E402
from dvt._lib import check, lib, ptr  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)


def make(M, N, K, bmn):
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(K, N, device="cuda", generator=g) if bmn else torch.randn(N, K, device="cuda", generator=g)
    ap, bp = ops.split_tf32(a), ops.split_tf32(b)
    out = torch.empty(M, N, device="cuda")

    def run(stream):
        check(lib().dvt_gemm_f32x3(ptr(ap), a.shape[1], a.numel(), 0, ptr(bp), b.shape[1], b.numel(), int(bmn), M, N, K, ptr(out), N, 1,
                                   None, stream.cuda_stream))
    return run


def timed(fns, reps=50):
    streams = [torch.cuda.Stream() for _ in fns]
    for f, s in zip(fns, streams):
        f(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_stream(main)
    for _ in range(reps):
        for f, s in zip(fns, streams):
            f(s)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


big = make(2048, 384, 768, True)     # data gradient of the field MLP: 96 CTAs, 24 k-blocks
small = make(1024, 384, 768, True)   # half of it: 48 CTAs
src = torch.empty(64 * 1024 * 1024, device="cuda")
dst = torch.empty_like(src)
copy = lambda s: dst.copy_(src) if s is None else _copy(s)  # noqa: E731


def _copy(s):
    with torch.cuda.stream(s):
        dst[: 8 * 1024 * 1024].copy_(src[: 8 * 1024 * 1024])  # 32 MB read + 32 MB written per call (~10 us at the HBM rate)


ta, tb = timed([big]), timed([small])
tab = timed([big, small])
tc = timed([_copy])
tac = timed([big, _copy])
print(f"96-CTA GEMM alone {ta:6.1f} us | 48-CTA GEMM alone {tb:6.1f} us | both side by side {tab:6.1f} us per pair")
print(f"copy alone {tc:6.1f} us | 96-CTA GEMM beside the copy {tac:6.1f} us per pair")
