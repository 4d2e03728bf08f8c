"""Unit operators over torch CUDA tensors (thin wrappers around the C ABI; used by tests and by dvt.models)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import DT_BF16, DT_F32, check, cur_stream, lib, ptr

_ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float32:
        return DT_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.DvtError("dvt_b200 operators need CUDA tensors (no CPU fallback)")


def gemm_tn(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act: str | None = None,
            out_dtype: torch.dtype = torch.bfloat16, splits: int = 1, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = act(a @ w.T + bias); a [M,K], w [N,K] (nn.Linear weight layout)."""
    _need_cuda(a, w, bias)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.dtype == w.dtype
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = (torch.zeros if splits > 1 else torch.empty)((M, N), device=a.device, dtype=out_dtype)
    check(lib().dvt_gemm_tn(ptr(a), a.stride(0), ptr(w), w.stride(0), _dt(a), M, N, K, ptr(bias), _ACT[act], ptr(out),
                            out.stride(0), _dt(out), splits, cur_stream()), "dvt_gemm_tn")
    return out


def gemm_tn_residual_(x: torch.Tensor, a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None,
                      gamma: torch.Tensor | None) -> torch.Tensor:
    """x += gamma * (a @ w.T + bias), x fp32 [M,N] in place."""
    _need_cuda(x, a, w, bias, gamma)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    check(lib().dvt_gemm_tn_residual(ptr(a), a.stride(0), ptr(w), w.stride(0), _dt(a), M, N, K, ptr(bias), ptr(gamma),
                                     ptr(x), x.stride(0), cur_stream()), "dvt_gemm_tn_residual")
    return x
