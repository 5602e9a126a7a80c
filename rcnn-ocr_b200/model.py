"""Host-side mirror of the reference's encoder modules (model/model.py) over the C ABI.

``BidirectionalLSTM(input_size, hidden_size, output_size)`` keeps the reference's parameter
names and shapes (``rnn.weight_ih_l0`` ... ``rnn.bias_hh_l0_reverse``, ``linear.weight``,
``linear.bias``; torch's i,f,g,o gate packing), so ``load_state_dict(strict=True)`` of the
encoder part of reference checkpoints works (model/model.py:151-163, training/utils.py:116).
Forward and backward run entirely in the sm_100a kernels:

    x --cast--> bf16 --K1 GEMM (W_ih, both directions)--> xp --K2 recurrence--> hcat (bf16)
      --K1 GEMM (linear)--> out                                         [forward]
    dout --K1 GEMMs--> dhcat, dW_lin --K2 BPTT--> dG --K1 GEMMs--> dX, dW_ih, dW_hh, db   [backward]

MMA operands are bf16 with fp32 accumulation; the cell state and gate pre-activations stay
fp32 (north_star tolerance: 1e-2 absolute on outputs against the fp32/fp64 reference path).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib, ops

_SUPPORTED_H = (64, 128, 256, 512)


def _cast2d(w: torch.Tensor) -> torch.Tensor:
    return ops.cast_bf16_3d(w.detach().unsqueeze(0))[0]


class _BiLSTMBlockFn(torch.autograd.Function):
    """One BidirectionalLSTM block: nn.LSTM(bidirectional, batch_first) + nn.Linear."""

    @staticmethod
    def forward(ctx, x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, lin_w, lin_b,
                out_dtype, save):
        _lib.require_cuda(x, "x")
        if x.dim() != 3:
            raise RuntimeError(f"BidirectionalLSTM expects [B, T, input_size], got {tuple(x.shape)}")
        B, T, I = x.shape
        H = w_hh_f.shape[1]
        O = lin_w.shape[0]
        packed = ops.lstm_pack(w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r)
        xb = ops.cast_bf16_3d(x)
        xp = ops.gemm_bf16(xb.view(B * T, I), packed.wih_p, packed.bias_p, torch.float32)
        hcat, gates, csave = ops.lstm_forward(xp, packed, B, T, save)
        del xp
        lin_wb = _cast2d(lin_w)
        out = ops.gemm_bf16(hcat.view(B * T, 2 * H), lin_wb, lin_b.detach().float().contiguous(), out_dtype)
        if save:
            ctx.save_for_backward(xb, hcat, gates, csave, lin_wb)
            ctx.packed = packed
            ctx.dims = (B, T, I, H, O)
            ctx.x_dtype = x.dtype
        return out.view(B, T, O)

    @staticmethod
    def backward(ctx, dout):
        xb, hcat, gates, csave, lin_wb = ctx.saved_tensors
        packed = ctx.packed
        B, T, I, H, O = ctx.dims
        BT = B * T
        dob = ops.cast_bf16_3d(dout)                                   # [B,T,O] bf16
        dob2 = dob.view(BT, O)
        # ---- linear: dhcat = dout W, dW = dout^T hcat, db = colsum(dout) --------------------
        lin_wt = ops.transpose_bf16(lin_wb)                            # [2H, O]
        dhcat = ops.gemm_bf16(dob2, lin_wt, None, torch.float32).view(B, T, 2 * H)
        dob_t = ops.transpose_bf16(dob2)                               # [O, BT]
        hcat_t = ops.transpose_bf16(hcat.view(BT, 2 * H))              # [2H, BT]
        d_lin_w = ops.gemm_bf16(dob_t, hcat_t, None, torch.float32)    # [O, 2H]
        d_lin_b = ops.colsum_bf16(dob2)
        del dob_t, hcat_t
        # ---- recurrence ---------------------------------------------------------------------
        dG = ops.lstm_backward(packed, gates, csave, dhcat, B, T)      # [B,T,8H] bf16
        dG2 = dG.view(BT, 8 * H)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm_bf16(dG2, packed.wih_pt, None, torch.float32).view(B, T, I).to(ctx.x_dtype)
        dG_t = ops.transpose_bf16(dG2)                                 # [8H, BT]
        xb_t = ops.transpose_bf16(xb.view(BT, I))                      # [I, BT]
        dwih_p = ops.gemm_bf16(dG_t, xb_t, None, torch.float32)        # [8H, I]
        hprev_t = ops.lstm_hprev_t(hcat)                               # [2, H, BT]
        dwhh_p = torch.empty((8 * H, H), dtype=torch.float32, device=dG.device)
        for d in range(2):
            ops.gemm_bf16(dG_t[d * 4 * H:(d + 1) * 4 * H], hprev_t[d], None, out=dwhh_p[d * 4 * H:(d + 1) * 4 * H])
        db_p = ops.colsum_bf16(dG2)
        g = ops.lstm_unpack_grads(dwih_p, dwhh_p, db_p, I, H)
        return (dx, g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], d_lin_w, d_lin_b, None, None)


class _LSTMParameters(nn.Module):
    """Parameter holder with nn.LSTM's names, shapes, registration order and init
    (U(-1/sqrt(H), 1/sqrt(H)) in registration order, so equal seeds give equal weights)."""

    def __init__(self, input_size: int, hidden_size: int):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        for sfx in ("", "_reverse"):
            self.register_parameter("weight_ih_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size, input_size)))
            self.register_parameter("weight_hh_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
            self.register_parameter("bias_ih_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size)))
            self.register_parameter("bias_hh_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size)))
        bound = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0.0
        for w in self.parameters():
            nn.init.uniform_(w, -bound, bound)

    def flatten_parameters(self):  # cuDNN-specific in the reference (model/model.py:160); nothing to do here
        return None

    def ordered(self):
        return [getattr(self, n + sfx) for sfx in ("", "_reverse")
                for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]


class BidirectionalLSTM(nn.Module):
    """Drop-in for model.model.BidirectionalLSTM (model/model.py:151-163).

    ``forward(x[B,T,input_size]) -> [B,T,output_size]``; accepts non-contiguous batch-first
    views (the encoder feeds a permuted [B,C,W] tensor, model/model.py:218).  hidden_size must
    be 64, 128, 256 or 512 and input_size / output_size multiples of 8."""

    def __init__(self, input_size: int, hidden_size: int, output_size: int, out_dtype: torch.dtype = torch.float32):
        super().__init__()
        if hidden_size not in _SUPPORTED_H:
            raise ValueError(f"hidden_size must be one of {_SUPPORTED_H}, got {hidden_size}")
        if input_size % 8 or output_size <= 0:
            raise ValueError("input_size must be a multiple of 8 (16-byte bf16 rows)")
        self.rnn = _LSTMParameters(input_size, hidden_size)
        self.linear = nn.Linear(hidden_size * 2, output_size)
        self.out_dtype = out_dtype

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        return _BiLSTMBlockFn.apply(x, *self.rnn.ordered(), self.linear.weight, self.linear.bias,
                                    self.out_dtype, save)


def make_enc_rnn(enc_dim: int, hidden_size: int) -> nn.Sequential:
    """RCNN.enc_rnn (model/model.py:195-198): two stacked blocks, state-dict keys '0.*', '1.*'."""
    return nn.Sequential(BidirectionalLSTM(enc_dim, hidden_size, hidden_size),
                         BidirectionalLSTM(hidden_size, hidden_size, hidden_size))
