"""Thin tensor-level wrappers over the C ABI (no autograd; see model.py for the modules)."""
from __future__ import annotations

import torch

from . import _lib

_OUT = {torch.float32: 0, torch.bfloat16: 1}


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None,
              out_dtype: torch.dtype = torch.float32, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T (+ bias[N]) on tcgen05 (kernel K1).  a, b: bf16, rows
    contiguous, row pitch a multiple of 8 elements."""
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "gemm_bf16 takes bf16 operands"
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1], (a.shape, b.shape)
    if a.stride(1) != 1 or a.stride(0) % 8 or a.data_ptr() % 16:
        a = a.contiguous()
    if b.stride(1) != 1 or b.stride(0) % 8 or b.data_ptr() % 16:
        b = b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    if K % 8:
        raise ValueError(f"K={K} must be a multiple of 8 (16-byte rows)")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in _OUT
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    with torch.cuda.device(a.device):
        rc = _lib.lib().rcnn_gemm_bf16(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                       out.stride(0), _OUT[out.dtype], bias.data_ptr() if bias is not None else None,
                                       M, N, K, _lib.stream_ptr())
        _lib.check(rc, "rcnn_gemm_bf16")
    return out
