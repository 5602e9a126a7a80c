"""Thin tensor-level wrappers over the C ABI (no autograd; see model.py for the modules)."""
from __future__ import annotations

import os

import torch

from . import _lib

_OUT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


# RCNN_POISON=1 (debugging aid): fill the buffers the recurrent kernels exchange data through (hcat, the backward
# workspace) with NaN patterns before the launch, so that a read that overtakes its write shows up as NaN
# instead of as the previous call's (usually identical) values.  scripts/stress_*.py
_POISON = os.environ.get("RCNN_POISON", "0") == "1"


def _repitch(m: torch.Tensor) -> torch.Tensor:
    """Copy a bf16 matrix into rows whose pitch is a multiple of 8 elements (host-side layout fix
    for oddly shaped operands; the hot path never takes this branch)."""
    R, C = m.shape
    buf = torch.zeros((R, (C + 7) // 8 * 8), dtype=m.dtype, device=m.device)
    buf[:, :C] = m
    return buf[:, :C]


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None,
              out_dtype: torch.dtype = torch.float32, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T (+ bias[N]) on tcgen05 (kernel K1).  a, b: bf16, rows
    contiguous, row pitch a multiple of 8 elements."""
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "gemm_bf16 takes bf16 operands"
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1], (a.shape, b.shape)
    if a.stride(1) != 1 or a.stride(0) % 8 or a.data_ptr() % 16:
        a = _repitch(a)
    if b.stride(1) != 1 or b.stride(0) % 8 or b.data_ptr() % 16:
        b = _repitch(b)
    M, K = a.shape
    N = b.shape[0]
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


def gemm_bf16_atb(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None,
                  accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= a[K,M]^T @ b[K,N] in fp32 (weight-gradient shape of kernel K1; the contraction
    index is the row of both row-major bf16 operands, so nothing is transposed)."""
    _lib.require_cuda(a, "a")
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.shape[0] == b.shape[0], (a.shape, b.shape)
    if a.stride(1) != 1 or a.stride(0) % 8 or a.data_ptr() % 16:
        a = _repitch(a)
    if b.stride(1) != 1 or b.stride(0) % 8 or b.data_ptr() % 16:
        b = _repitch(b)
    K, M = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        accumulate = False
    assert out.shape == (M, N) and out.dtype == torch.float32 and out.stride(1) == 1
    with torch.cuda.device(a.device):
        rc = _lib.lib().rcnn_gemm_bf16_atb(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                           out.stride(0), M, N, K, int(accumulate), _lib.stream_ptr())
        _lib.check(rc, "rcnn_gemm_bf16_atb")
    return out


def gemm_bf16_atb_grouped(a: torch.Tensor, b: torch.Tensor, groups: int, M: int, N: int) -> torch.Tensor:
    """out[g] = a[:, g*M:(g+1)*M]^T @ b[:, g*N:(g+1)*N] for g < groups, one launch; returns [groups*M, N] fp32."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape[0] == b.shape[0]
    assert a.shape[1] == groups * M and b.shape[1] == groups * N and a.stride(1) == 1 and b.stride(1) == 1
    K = a.shape[0]
    with torch.cuda.device(a.device):
        out = torch.empty((groups * M, N), dtype=torch.float32, device=a.device)
        rc = _lib.lib().rcnn_gemm_bf16_atb_grouped(a.data_ptr(), a.stride(0), M, b.data_ptr(), b.stride(0), N,
                                                   out.data_ptr(), N, M * N, groups, M, N, K, 0, _lib.stream_ptr())
        _lib.check(rc, "rcnn_gemm_bf16_atb_grouped")
    return out


class PackedLSTMWeights:
    """bf16 views of one block's nn.LSTM parameters in the kernels' layouts (see
    rcnn_lstm_pack_weights in include/rcnn_ocr_b200.h)."""

    def __init__(self, blob: torch.Tensor, I: int, H: int):
        self.blob, self.I, self.H = blob, I, H
        H8 = 8 * H
        off = 0

        def take(nbytes, dtype, shape):
            nonlocal off
            v = blob[off: off + nbytes].view(dtype).view(shape)
            off += nbytes
            return v

        self.wih_p = take(H8 * I * 2, torch.bfloat16, (H8, I))
        self.bias_p = take(H8 * 4, torch.float32, (H8,))
        self.whh_p = take(H8 * H * 2, torch.bfloat16, (H8, H))
        self.whh_pt = take(H8 * H * 2, torch.bfloat16, (2, H, 4 * H))
        self.wih_pt = take(H8 * I * 2, torch.bfloat16, (I, H8))


def lstm_pack(w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, parts: int = 3,
              into: PackedLSTMWeights | None = None) -> PackedLSTMWeights:
    """parts: 1 = forward views, 2 = transposed (backward) views, 3 = both; ``into`` fills an existing blob
    (the two halves may be converted on different streams)."""
    ts = [w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r]
    for t in ts:
        _lib.require_cuda(t, "LSTM parameter")
    ts = [t.detach().float().contiguous() for t in ts]
    H = ts[1].shape[1]
    I = ts[0].shape[1]
    L = _lib.lib()
    nbytes = L.rcnn_lstm_packed_bytes(I, H)
    with torch.cuda.device(ts[0].device):
        blob = into.blob if into is not None else torch.empty((nbytes,), dtype=torch.uint8, device=ts[0].device)
        rc = L.rcnn_lstm_pack_weights_parts(*[t.data_ptr() for t in ts], I, H, blob.data_ptr(), parts, _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_pack_weights")
    return into if into is not None else PackedLSTMWeights(blob, I, H)


def cast_bf16_3d(x: torch.Tensor) -> torch.Tensor:
    """[B,T,C] float32 in any strides -> contiguous bf16 (bf16 contiguous inputs pass through)."""
    _lib.require_cuda(x, "x")
    if x.dtype == torch.bfloat16:
        return x if x.is_contiguous() else x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    B, T, C = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, T, C), dtype=torch.bfloat16, device=x.device)
        rc = _lib.lib().rcnn_cast_bf16_3d(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), out.data_ptr(),
                                          B, T, C, _lib.stream_ptr())
        _lib.check(rc, "rcnn_cast_bf16_3d")
    return out


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    """[R,C] bf16 (row stride arbitrary) -> [C,R] whose row pitch is padded to a multiple of 8
    elements (16-byte rows for TMA); the pad columns are never read."""
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    ldo = _pad8(R)
    with torch.cuda.device(x.device):
        out = torch.empty((C, ldo), dtype=torch.bfloat16, device=x.device)
        rc = _lib.lib().rcnn_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), ldo, R, C, _lib.stream_ptr())
        _lib.check(rc, "rcnn_transpose_bf16")
    return out[:, :R]


def lstm_forward(xp: torch.Tensor, packed: PackedLSTMWeights, B: int, T: int, save: bool, hcat=None):
    """Recurrence of both directions (kernel K2).  xp: float16 [B*T, 8H] in packed column order.
    Returns (hcat bf16 [B,T,2H], gates f16 [2,T,B,4H] | None, c f32 [2,T,B,H] | None)."""
    H = packed.H
    assert xp.dtype == torch.float16 and xp.shape == (B * T, 8 * H) and xp.is_contiguous()
    dev = xp.device
    with torch.cuda.device(dev):
        if hcat is None:
            hcat = torch.empty((B, T, 2 * H), dtype=torch.bfloat16, device=dev)
            if _POISON:
                hcat.view(torch.int16).fill_(-1)
        gates = torch.empty((2, T, B, 4 * H), dtype=torch.float16, device=dev) if save else None
        csave = torch.empty((2, T, B, H), dtype=torch.float32, device=dev) if save else None
        rc = _lib.lib().rcnn_lstm_forward(xp.data_ptr(), packed.whh_p.data_ptr(), B, T, H, hcat.data_ptr(),
                                          gates.data_ptr() if save else None, csave.data_ptr() if save else None,
                                          _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_forward")
    return hcat, gates, csave


def lstm_forward_fused(xb: torch.Tensor, packed: PackedLSTMWeights, B: int, T: int, save: bool = False):
    """Recurrence with the input projection fused in (no xp tensor).  xb: bf16 [B, T, I] contiguous, I a
    multiple of 64 and <= 512.  Returns (hcat bf16 [B,T,2H], gates f16 [2,T,B,4H] | None, c f32 [2,T,B,H] | None)."""
    H, I = packed.H, packed.I
    assert xb.dtype == torch.bfloat16 and xb.shape == (B, T, I) and xb.is_contiguous()
    dev = xb.device
    with torch.cuda.device(dev):
        hcat = torch.empty((B, T, 2 * H), dtype=torch.bfloat16, device=dev)
        if _POISON:
            hcat.view(torch.int16).fill_(-1)
        gates = torch.empty((2, T, B, 4 * H), dtype=torch.float16, device=dev) if save else None
        csave = torch.empty((2, T, B, H), dtype=torch.float32, device=dev) if save else None
        rc = _lib.lib().rcnn_lstm_forward_fused(xb.data_ptr(), packed.wih_p.data_ptr(), packed.bias_p.data_ptr(),
                                                packed.whh_p.data_ptr(), B, T, I, H, hcat.data_ptr(),
                                                gates.data_ptr() if save else None, csave.data_ptr() if save else None,
                                                _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_forward_fused")
    return hcat, gates, csave


def fused_forward_supported(I: int, H: int) -> bool:
    return I % 64 == 0 and 64 <= I <= 512 and H in (64, 128, 256, 512)


def lstm_backward(packed: PackedLSTMWeights, gates, csave, dhcat, B: int, T: int):
    """BPTT of both directions (kernel K2 backward).  dhcat: float32 [B,T,2H] contiguous.
    Returns (dG bf16 [B,T,8H]: gradient w.r.t. the gate pre-activations, packed column order;
    db f32 [8H]: its column sums = the packed-order bias gradient, accumulated inside the kernel)."""
    H = packed.H
    assert dhcat.dtype == torch.float32 and dhcat.is_contiguous() and dhcat.shape == (B, T, 2 * H)
    with torch.cuda.device(dhcat.device):
        dG = torch.empty((B, T, 8 * H), dtype=torch.bfloat16, device=dhcat.device)
        db = torch.empty((8 * H,), dtype=torch.float32, device=dhcat.device)
        nws = int(_lib.lib().rcnn_lstm_backward_workspace_bytes(B, T, H))
        ws = torch.empty((max(nws, 1),), dtype=torch.uint8, device=dhcat.device)
        if _POISON:
            ws.fill_(0xFF)
        rc = _lib.lib().rcnn_lstm_backward(packed.whh_pt.data_ptr(), gates.data_ptr(), csave.data_ptr(),
                                           dhcat.data_ptr(), B, T, H, dG.data_ptr(), db.data_ptr(), ws.data_ptr(), nws,
                                           _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_backward")
    return dG, db


def lstm_plan(B: int, H: int, backward: bool = False):
    """(work items a CTA group works on at a time, CTA groups) of the recurrent kernels for a batch of B sequences:
    two interleaved items per group once items would otherwise queue behind each other (rcnn_lstm_plan)."""
    import ctypes
    nslot, ngroups = ctypes.c_int(0), ctypes.c_int(0)
    rc = _lib.lib().rcnn_lstm_plan(int(backward), B, H, ctypes.byref(nslot), ctypes.byref(ngroups))
    _lib.check(rc, "rcnn_lstm_plan")
    return nslot.value, ngroups.value


def cast_bf16_2d(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """float32 [R,C] (rows contiguous) -> bf16 [R,C] view whose row pitch is padded to a multiple of
    8 elements (pad columns zero), ready to be a GEMM operand for any C.  ``out``: an existing bf16 buffer
    [>= R, pad8(C)] whose first R rows are written."""
    _lib.require_cuda(x, "x")
    assert x.dim() == 2
    if x.dtype != torch.float32:
        x = x.float()
    if x.shape[1] > 1 and x.stride(1) != 1:
        x = x.contiguous()
    R, C = x.shape
    ldd = _pad8(C)
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty((R, ldd), dtype=torch.bfloat16, device=x.device)
        else:
            assert out.dtype == torch.bfloat16 and out.shape[0] >= R and out.shape[1] == ldd and out.is_contiguous()
        rc = _lib.lib().rcnn_cast_bf16_2d(x.data_ptr(), x.stride(0), out.data_ptr(), ldd, R, C, _lib.stream_ptr())
        _lib.check(rc, "rcnn_cast_bf16_2d")
    return out[:R, :C]


def colsum_bf16(x: torch.Tensor) -> torch.Tensor:
    """float32 column sums of a bf16 [rows, cols] matrix (rows contiguous, any row pitch)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and (x.stride(1) == 1 or x.shape[1] <= 1)
    with torch.cuda.device(x.device):
        out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
        rc = _lib.lib().rcnn_colsum_bf16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "rcnn_colsum_bf16")
    return out


def lstm_hprev(hcat: torch.Tensor) -> torch.Tensor:
    """[B,T,2H] bf16 -> [B,T,2H]: each direction's previous-step h (zero at its first step)."""
    B, T, H2 = hcat.shape
    assert hcat.is_contiguous() and hcat.dtype == torch.bfloat16
    with torch.cuda.device(hcat.device):
        out = torch.empty_like(hcat)
        rc = _lib.lib().rcnn_lstm_hprev(hcat.data_ptr(), out.data_ptr(), B, T, H2 // 2, _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_hprev")
    return out


def lstm_unpack_grads(dwih_p, dwhh_p, db_p, I: int, H: int):
    """Packed-order weight gradients -> eight float32 tensors in torch's nn.LSTM layout
    (w_ih, w_hh, b_ih, b_hh for the forward direction, then the reverse one)."""
    dev = dwih_p.device
    with torch.cuda.device(dev):
        outs = []
        for _ in range(2):
            outs += [torch.empty((4 * H, I), dtype=torch.float32, device=dev),
                     torch.empty((4 * H, H), dtype=torch.float32, device=dev),
                     torch.empty((4 * H,), dtype=torch.float32, device=dev),
                     torch.empty((4 * H,), dtype=torch.float32, device=dev)]
        rc = _lib.lib().rcnn_lstm_unpack_grads(dwih_p.data_ptr(), dwhh_p.data_ptr(), db_p.data_ptr(), I, H,
                                               *[o.data_ptr() for o in outs], _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_unpack_grads")
    return outs


def weight_grads_supported(I: int, H: int, T: int) -> bool:
    """rcnn_lstm_weight_grads: the CTA-pair shapes; K chunks are 64 steps of ONE sequence, so T should fill them."""
    return H in (256, 512) and I >= 256 and I % 32 == 0 and (-(-T // 64) * 64) <= 1.25 * T


def lstm_weight_grads(dG: torch.Tensor, xb: torch.Tensor, hcat: torch.Tensor, db_p: torch.Tensor, B: int, T: int, I: int, H: int):
    """dW_ih, dW_hh, db of one block straight in nn.LSTM's layout: two GEMM launches (h_{t-/+1} read from hcat
    through a shifted tensor map, rows scattered to gate-major order by the reduce-add map) and one small copy
    for the biases.  Returns the eight gradients in ``_LSTMParameters.ordered()`` order."""
    assert dG.dtype == torch.bfloat16 and dG.is_contiguous() and dG.numel() == B * T * 8 * H
    assert xb.dtype == torch.bfloat16 and xb.is_contiguous() and hcat.dtype == torch.bfloat16 and hcat.is_contiguous()
    dev = dG.device
    with torch.cuda.device(dev):
        n_ih, n_hh = 2 * 4 * H * I, 2 * 4 * H * H
        buf = torch.zeros((n_ih + n_hh,), dtype=torch.float32, device=dev)        # one fill for both outputs
        dwih, dwhh = buf[:n_ih].view(2, 4 * H, I), buf[n_ih:].view(2, 4 * H, H)
        rc = _lib.lib().rcnn_lstm_weight_grads(dG.data_ptr(), xb.data_ptr(), hcat.data_ptr(), B, T, I, H,
                                               dwih.data_ptr(), dwhh.data_ptr(), 1, _lib.stream_ptr())
        _lib.check(rc, "rcnn_lstm_weight_grads")
        # packed bias order (unit-major, gate-minor) -> gate-major, one copy each for b_ih and b_hh (same values)
        db = db_p.view(2, 1, H, 4).permute(0, 1, 3, 2).expand(2, 2, 4, H).reshape(2, 2, 4 * H)
    return [dwih[0], dwhh[0], db[0, 0], db[0, 1], dwih[1], dwhh[1], db[1, 0], db[1, 1]]


def launch_count() -> int:
    """Kernels launched by the library so far in this process."""
    return int(_lib.lib().rcnn_launch_count())
