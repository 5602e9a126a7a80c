"""Host-side mirror of the reference's encoder modules (model/model.py) over the C ABI.

``BidirectionalLSTM(input_size, hidden_size, output_size)`` keeps the reference's parameter
names and shapes (``rnn.weight_ih_l0`` ... ``rnn.bias_hh_l0_reverse``, ``linear.weight``,
``linear.bias``; torch's i,f,g,o gate packing), so ``load_state_dict(strict=True)`` of the
encoder part of reference checkpoints works (model/model.py:151-163, training/utils.py:116).
Forward and backward run entirely in the sm_100a kernels:

    x --cast--> bf16 --K1 GEMM (W_ih, both directions)--> xp (fp16) --K2 recurrence--> hcat (bf16)
      --K1 GEMM (linear)--> out                                         [forward]
    dout --K1 GEMMs--> dhcat, dW_lin --K2 BPTT--> dG --K1 GEMMs--> dX, dW_ih, dW_hh, db   [backward]

MMA operands are bf16 with fp32 accumulation; the cell state and gate pre-activations stay
fp32 (north_star tolerance: 1e-2 absolute on outputs against the fp32/fp64 reference path).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib, ops

_SUPPORTED_H = (64, 128, 256, 512)


def _cast2d(w: torch.Tensor) -> torch.Tensor:
    return ops.cast_bf16_3d(w.detach().unsqueeze(0))[0]


class _Prepared:
    """bf16 kernel views of one block's parameters.  ``rows_event`` / ``rest_event``: when the views were converted
    on the prefetch stream (EncRNN.forward), the events the consuming stream waits on before the recurrence
    (wih_p, bias_p, whh_p) and before the Linear / the backward pass (lin_wb, lin_wt, whh_pt, wih_pt)."""
    __slots__ = ("packed", "lin_wb", "lin_wt", "rows_event", "rest_event")

    def __init__(self, packed, lin_wb, lin_wt=None, rows_event=None, rest_event=None):
        self.packed, self.lin_wb, self.lin_wt = packed, lin_wb, lin_wt
        self.rows_event, self.rest_event = rows_event, rest_event

    def wait_rows(self):
        if self.rows_event is not None:
            torch.cuda.current_stream().wait_event(self.rows_event)
            self.rows_event = None

    def wait_rest(self):
        if self.rest_event is not None:
            torch.cuda.current_stream().wait_event(self.rest_event)
            self.rest_event = None


_PREFETCH = os.environ.get("RCNN_PREFETCH", "1") != "0"
_side_streams = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    key = (device.index if device.index is not None else torch.cuda.current_device())
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class _BiLSTMBlockFn(torch.autograd.Function):
    """One BidirectionalLSTM block: nn.LSTM(bidirectional, batch_first) + nn.Linear."""

    @staticmethod
    def forward(ctx, x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, lin_w, lin_b,
                out_dtype, save, prepared=None):
        _lib.require_cuda(x, "x")
        if x.dim() != 3:
            raise RuntimeError(f"BidirectionalLSTM expects [B, T, input_size], got {tuple(x.shape)}")
        B, T, I = x.shape
        H = w_hh_f.shape[1]
        O = lin_w.shape[0]
        if prepared is None:    # bf16 kernel views of the parameters (cached by the module in eval mode)
            prepared = _Prepared(ops.lstm_pack(w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r), _cast2d(lin_w))
        packed, lin_wb = prepared.packed, prepared.lin_wb
        xb = ops.cast_bf16_3d(x)
        prepared.wait_rows()
        if ops.fused_forward_supported(I, H):
            # the input projection runs inside the recurrent kernel (no xp tensor)
            hcat, gates, csave = ops.lstm_forward_fused(xb, packed, B, T, save)
        else:
            xp = ops.gemm_bf16(xb.view(B * T, I), packed.wih_p, packed.bias_p, torch.float16)
            hcat, gates, csave = ops.lstm_forward(xp, packed, B, T, save)
            del xp
        prepared.wait_rest()
        out = ops.gemm_bf16(hcat.view(B * T, 2 * H), lin_wb, lin_b.detach().float().contiguous(), out_dtype)
        if save:
            ctx.save_for_backward(xb, hcat, gates, csave, lin_wb)
            ctx.packed = packed
            ctx.lin_wt = prepared.lin_wt
            ctx.dims = (B, T, I, H, O)
            ctx.x_dtype = x.dtype
            # the parameter objects whose gradients this function returns (for dist.mark_grads_ready)
            ctx.params = [t for t in (w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, lin_w, lin_b)
                          if isinstance(t, torch.nn.Parameter)] or None
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
        lin_wt = ctx.lin_wt if ctx.lin_wt is not None else ops.transpose_bf16(lin_wb)   # [2H, O] (weights only: small)
        dhcat = ops.gemm_bf16(dob2, lin_wt, None, torch.float32).view(B, T, 2 * H)
        d_lin_w = ops.gemm_bf16_atb(dob2, hcat.view(BT, 2 * H))        # [O, 2H] = dout^T hcat
        d_lin_b = ops.colsum_bf16(dob2)
        # ---- recurrence ---------------------------------------------------------------------
        dG, db_p = ops.lstm_backward(packed, gates, csave, dhcat, B, T)   # [B,T,8H] bf16, [8H] f32
        dG2 = dG.view(BT, 8 * H)
        # weight gradients first: under data parallelism their all-reduce starts at the event recorded here and runs
        # while the input-gradient GEMM below (and whatever backward work precedes this block in the model) executes
        if ops.weight_grads_supported(I, H, T):
            g = ops.lstm_weight_grads(dG, xb, hcat, db_p, B, T, I, H)  # torch layout, no h_prev copy / unpack pass
        else:
            dwih_p = ops.gemm_bf16_atb(dG2, xb.view(BT, I))                # [8H, I] = dG^T x
            hprev = ops.lstm_hprev(hcat).view(BT, 2 * H)
            dwhh_p = ops.gemm_bf16_atb_grouped(dG2, hprev, 2, 4 * H, H)   # per direction: [4H, H] = dG_d^T h_prev_d
            g = ops.lstm_unpack_grads(dwih_p, dwhh_p, db_p, I, H)
        if ctx.params is not None:
            from .dist import mark_grads_ready
            mark_grads_ready(ctx.params)
        dx = None
        if ctx.needs_input_grad[0]:
            dx_dtype = torch.bfloat16 if ctx.x_dtype == torch.bfloat16 else torch.float32
            dx = ops.gemm_bf16(dG2, packed.wih_pt, None, dx_dtype).view(B, T, I).to(ctx.x_dtype)
        return (dx, g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], d_lin_w, d_lin_b, None, None, None)


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
        self._prepared = None     # (key, _Prepared) while the module is in eval mode
        self._prefetched = None   # _Prepared handed over by EncRNN.forward for the next call (training)

    def _prepared_weights(self):
        """eval() mode: the packed bf16 views are rebuilt only when a parameter was replaced or
        modified in place (storage pointer / version counter), not on every call."""
        ws = self.rnn.ordered() + [self.linear.weight]
        key = tuple((w.data_ptr(), w._version, w.device) for w in ws)
        if self._prepared is None or self._prepared[0] != key:
            with torch.no_grad():
                self._prepared = (key, _Prepared(ops.lstm_pack(*self.rnn.ordered()), _cast2d(self.linear.weight)))
        return self._prepared[1]

    def prefetch_weights(self) -> "_Prepared":
        """Training: convert this block's parameters for the kernels on the prefetch stream, concurrently with
        whatever the current stream runs next (the input cast; the recurrent kernels leave 20 SMs idle).  The
        forward views come first and have their own event."""
        main = torch.cuda.current_stream()
        side = _side_stream(self.linear.weight.device)
        with torch.no_grad():
            side.wait_stream(main)
            with torch.cuda.stream(side):
                packed = ops.lstm_pack(*self.rnn.ordered(), parts=1)
                ev_rows = side.record_event()
                ops.lstm_pack(*self.rnn.ordered(), parts=2, into=packed)
                lin_wb = _cast2d(self.linear.weight)
                lin_wt = ops.transpose_bf16(lin_wb)
                ev = side.record_event()
            for t in (packed.blob, lin_wb, lin_wt):
                t.record_stream(main)
        return _Prepared(packed, lin_wb, lin_wt, rows_event=ev_rows, rest_event=ev)

    def train(self, mode: bool = True):
        if mode:
            self._prepared = None
        return super().train(mode)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        prepared = self._prefetched
        self._prefetched = None
        if prepared is None and not self.training and x.is_cuda:
            prepared = self._prepared_weights()
        return _BiLSTMBlockFn.apply(x, *self.rnn.ordered(), self.linear.weight, self.linear.bias,
                                    self.out_dtype, save, prepared)


class EncRNN(nn.Sequential):
    """nn.Sequential of BidirectionalLSTM blocks (same state-dict keys); in training the parameter conversions of
    all blocks run on a second stream, under the input cast and the first recurrence."""

    def forward(self, x):
        if _PREFETCH and self.training and x.is_cuda and torch.is_grad_enabled():
            for blk in self:
                blk._prefetched = blk.prefetch_weights()
        return super().forward(x)


def make_enc_rnn(enc_dim: int, hidden_size: int, out_dtype: torch.dtype = torch.float32) -> nn.Sequential:
    """RCNN.enc_rnn (model/model.py:195-198): two stacked blocks, state-dict keys '0.*', '1.*'."""
    # the first block hands bf16 straight to the second (which would cast its input anyway); the
    # stack's output stays fp32 like the reference's unless the caller feeds it to another bf16
    # GEMM (the CTC head) and asks for the hand-off format: same rounding, one cast kernel fewer
    return EncRNN(BidirectionalLSTM(enc_dim, hidden_size, hidden_size, out_dtype=torch.bfloat16),
                  BidirectionalLSTM(hidden_size, hidden_size, hidden_size, out_dtype=out_dtype))


_PAD_CACHE = {}   # (weight pointer, N, K, device) -> (zero-padded bf16 weight buffer, zero-padded bias buffer), see _LinearFn


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (bf16 operands, fp32 accumulate), any N."""

    @staticmethod
    def forward(ctx, x, weight, bias, save):
        _lib.require_cuda(x, "x")
        lead = x.shape[:-1]
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        xb = x2 if x2.dtype == torch.bfloat16 and x2.is_contiguous() else ops.cast_bf16_2d(x2)
        N, M = weight.shape[0], x2.shape[0]
        if N % 32 and N > 128 and M >= 512:
            # An odd class count (C = 195: rows of 780 bytes) keeps the GEMM off its CTA-pair kernel and off the TMA-store
            # epilogue (27 us for 3.3 GFLOP).  The weight is cast into the first N rows of a zero-padded [256k, K] buffer
            # (kept per weight tensor: the pad rows are written once), the product is computed N-padded with 16-byte
            # aligned output rows and the caller gets the [.., :N] view -- the CTC and decode kernels take a row pitch.
            Np = (N + 255) // 256 * 256
            key = (weight.data_ptr(), N, K, str(weight.device))
            pad = _PAD_CACHE.get(key)
            if pad is None:
                pad = (torch.zeros((Np, (K + 7) // 8 * 8), dtype=torch.bfloat16, device=weight.device),
                       torch.zeros((Np,), dtype=torch.float32, device=weight.device))
                _PAD_CACHE[key] = pad
            wb = ops.cast_bf16_2d(weight.detach(), out=pad[0])
            if bias is not None:
                pad[1][:N].copy_(bias.detach())
            out = ops.gemm_bf16(xb, pad[0][:, :K], pad[1] if bias is not None else None, torch.float32)[:, :N]
        else:
            wb = ops.cast_bf16_2d(weight.detach())
            out = ops.gemm_bf16(xb, wb, bias.detach().float().contiguous() if bias is not None else None, torch.float32)
        if save:
            ctx.save_for_backward(xb, wb)
            ctx.x_dtype = x.dtype
            ctx.has_bias = bias is not None
        return out.view(*lead, N) if out.is_contiguous() else out.unflatten(0, lead)

    @staticmethod
    def backward(ctx, dout):
        xb, wb = ctx.saved_tensors
        N, K = wb.shape
        M = xb.shape[0]
        dob = ops.cast_bf16_2d(dout.reshape(M, N))            # row pitch padded to 8 for odd N (C = 195)
        dx = None
        if ctx.needs_input_grad[0]:
            wt = ops.transpose_bf16(wb)                        # [K, N]
            dx_dtype = torch.bfloat16 if ctx.x_dtype == torch.bfloat16 else torch.float32
            dx = ops.gemm_bf16(dob, wt, None, dx_dtype).view(*dout.shape[:-1], K).to(ctx.x_dtype)
        dw = ops.gemm_bf16_atb(dob, xb)                         # [N, K] = dout^T x
        db = ops.colsum_bf16(dob) if ctx.has_bias else None
        return dx, dw, db, None


class CTCHead(nn.Module):
    """Linear(hidden -> num_ctc_classes) over enc[B,T,H] -> logits[B,T,C] (fp32).  Not present in
    the reference (its live head is the attention decoder); the north_star adds it behind the
    unchanged RCNN.forward signature."""

    def __init__(self, hidden_size: int, num_ctc_classes: int):
        super().__init__()
        lin = nn.Linear(hidden_size, num_ctc_classes)
        self.weight, self.bias = lin.weight, lin.bias

    def forward(self, enc: torch.Tensor) -> torch.Tensor:
        save = torch.is_grad_enabled() and (enc.requires_grad or self.weight.requires_grad)
        return _LinearFn.apply(enc, self.weight, self.bias, save)


# ---------------------------------------------------------------------------------------------
# Backbone adapter.  The SE-ResNet31 conv stack is outside the hot path (SURVEY.md section 2 #5:
# it stays on cuDNN through PyTorch); it is restated here only so that RCNN keeps the reference's
# constructor, forward signature and state-dict keys (model/seresnet31.py:70-187).
# ---------------------------------------------------------------------------------------------

class _SEGate(nn.Module):
    def __init__(self, ch: int, reduction: int):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(ch, ch // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(ch // reduction, ch, bias=False), nn.Sigmoid())

    def forward(self, x):
        return x * self.fc(x.mean(dim=(2, 3)))[:, :, None, None]


class _SEResidual(nn.Module):
    def __init__(self, cin, cout, stride, reduction, dropblock_p, dropblock_block_size):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.se = _SEGate(cout, reduction)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        if dropblock_p > 0:
            from torchvision.ops import DropBlock2d
            self.dropblock = DropBlock2d(p=dropblock_p, block_size=dropblock_block_size)
        else:
            self.dropblock = nn.Identity()

    def forward(self, x):
        y = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        y = self.dropblock(self.se(y))
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(y + skip)


class SEResNet31(nn.Module):
    _STAGES = ((128, 256, 1, 2), (256, 256, 2, 1), (256, 512, 5, 2), (512, 512, 3, 1))  # cin, cout, blocks, stride

    def __init__(self, in_channels=3, out_channels=512, reduction=16, dropblock_p=0.0, dropblock_block_size=5):
        super().__init__()

        def cbr(ci, co, k, s, p):
            return [nn.Conv2d(ci, co, k, s, p, bias=False), nn.BatchNorm2d(co), nn.ReLU(True)]

        self.conv0 = nn.Sequential(*cbr(in_channels, 64, 3, 1, 1), *cbr(64, 128, 3, 1, 1), nn.MaxPool2d(2, 2))
        for idx, (ci, co, n, s) in enumerate(self._STAGES, 1):
            blocks = [_SEResidual(ci if i == 0 else co, co, s if i == 0 else 1, reduction, dropblock_p,
                                  dropblock_block_size) for i in range(n)]
            setattr(self, f"layer{idx}", nn.Sequential(*blocks))
        self.conv_out = nn.Sequential(*cbr(512, out_channels, 2, (2, 1), (0, 1)), *cbr(out_channels, out_channels, 2, 1, 0))
        self.out_channels = out_channels

    def forward(self, x):
        x = self.conv0(x)
        for idx in range(1, 5):
            x = getattr(self, f"layer{idx}")(x)
        return self.conv_out(x)


class FoldedBackbone(nn.Module):
    """Inference-only copy of an ``SEResNet31`` (the cheap wins): every BatchNorm folded into the
    convolution before it (eval statistics); conv + bias + ReLU as ONE cuDNN call where a ReLU follows directly
    (``torch.cudnn_convolution_relu``); the squeeze-and-excitation tail of every block (mean, two small products, sigmoid,
    scale, residual add, ReLU -- and the biases of the two convolutions that feed it) as two launches of this library
    (csrc/se_gate.cu) instead of ten; the stem's max pooling on this library's kernel; the 3-channel input padded to 8
    channels so that the first convolution takes cuDNN's tensor-core path; weights held in ``dtype`` (bf16) and
    channels_last.  About 130 fewer launches per image batch than the module it copies; the module itself (and its state
    dict) is left untouched.  Built from a snapshot of the weights: rebuild after loading a checkpoint.  One stream at a time:
    the blocks share the SE kernels' workspace (as the reference's single Python thread / default stream does)."""

    def __init__(self, cnn: "SEResNet31", dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        import copy
        from torch.nn.utils.fusion import fuse_conv_bn_weights
        self.dtype = dtype
        self._n = 0
        self._ws = {}                                        # (B, C, device) -> zeroed workspace of rcnn_se_gate

        def fold(conv: nn.Conv2d, bn: nn.BatchNorm2d, pad_in: int = 0):
            w, b = fuse_conv_bn_weights(conv.weight.detach().float(), None if conv.bias is None else conv.bias.detach().float(),
                                        bn.running_mean.float(), bn.running_var.float(), bn.eps,
                                        bn.weight.detach().float(), bn.bias.detach().float())
            if pad_in > w.shape[1]:                          # zero weights for the padded input channels
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_in - w.shape[1], *w.shape[2:])], 1)
            i = self._n
            self._n += 1
            self.register_buffer(f"w{i}", w.detach().to(dtype).contiguous(memory_format=torch.channels_last), persistent=False)
            self.register_buffer(f"b{i}", b.detach().to(dtype).contiguous(), persistent=False)
            self.register_buffer(f"bf{i}", b.detach().float().contiguous(), persistent=False)     # f32 copy for the fused tail
            return (i, tuple(conv.stride), tuple(conv.padding))

        def seq(mods, pad_first: int = 0):                   # [conv, bn, relu, conv, bn, relu(, pool)] -> folded conv + relu pairs
            mods = list(mods)
            out, k = [], 0
            while k < len(mods):
                if isinstance(mods[k], nn.Conv2d):
                    out.append(("cr", fold(mods[k], mods[k + 1], pad_first if not out else 0)))
                    k += 3
                else:
                    out.append(("pool", mods[k]))
                    k += 1
            return out

        self.in_channels = cnn.conv0[0].in_channels
        self.in_pad = 8 if self.in_channels < 8 else self.in_channels
        self.stem = seq(cnn.conv0, self.in_pad)
        self.tail = seq(cnn.conv_out)
        self.blocks = []
        self.se = nn.ModuleList()
        for idx in range(1, 5):
            for blk in getattr(cnn, f"layer{idx}"):
                ds = None if blk.downsample is None else fold(blk.downsample[0], blk.downsample[1])
                self.blocks.append((fold(blk.conv1, blk.bn1), fold(blk.conv2, blk.bn2), ds))
                self.se.append(copy.deepcopy(blk.se.fc).to(dtype))
                k = len(self.blocks) - 1                     # f32 copies of the two SE products for the fused tail (se_gate.cu)
                self.register_buffer(f"se{k}_w1", blk.se.fc[0].weight.detach().float().contiguous(), persistent=False)
                self.register_buffer(f"se{k}_w2", blk.se.fc[2].weight.detach().float().t().contiguous(), persistent=False)  # [Cr, C]
        self.pools = nn.ModuleList([m for kind, m in self.stem if kind == "pool"])

    def _conv(self, x, spec, relu: bool, bias: bool = True):
        i, stride, padding = spec
        w, b = getattr(self, f"w{i}"), (getattr(self, f"b{i}") if bias else None)
        if relu and x.is_cuda:
            return torch.cudnn_convolution_relu(x, w, b, stride, padding, (1, 1), 1)
        y = torch.nn.functional.conv2d(x, w, b, stride, padding)
        return torch.relu_(y) if relu else y

    @staticmethod
    def _fusable(t: torch.Tensor) -> bool:
        vec = 8 if t.dtype == torch.bfloat16 else 4
        return (t.is_cuda and t.dtype in (torch.bfloat16, torch.float32) and t.shape[1] % vec == 0
                and t.is_contiguous(memory_format=torch.channels_last))

    def _pool(self, pool, x):
        """nn.MaxPool2d(2, 2): this library's channels_last kernel when it applies (torch's ran at 0.9 TB/s)."""
        B, C, H, W = x.shape
        plain = (pool.kernel_size in (2, (2, 2)) and pool.stride in (2, (2, 2)) and pool.padding in (0, (0, 0))
                 and pool.dilation in (1, (1, 1)) and not pool.ceil_mode)
        if plain and self._fusable(x) and H % 2 == 0 and W % 2 == 0:
            from . import _lib
            with torch.cuda.device(x.device):
                out = torch.empty((B, C, H // 2, W // 2), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
                _lib.check(_lib.lib().rcnn_maxpool2x2_nhwc(x.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0, B, H, W, C,
                                                           out.data_ptr(), _lib.stream_ptr()), "rcnn_maxpool2x2_nhwc")
            return out
        return pool(x)

    def _se_tail(self, k, x, c2, ds, y, fc):
        """relu((y + b2) * sigmoid(W2 relu(W1 mean_hw(y + b2))) + skip): two launches of this library (csrc/se_gate.cu) on a
        CUDA device -- y is conv2's output WITHOUT its bias and the downsample convolution runs without its bias as well: both
        are added inside the kernels -- the torch ops otherwise (the adapter also runs on the host for the CPU baseline)."""
        B, C, H, W = y.shape
        skip = x if ds is None else self._conv(x, ds, False, bias=False)
        if self._fusable(y) and self._fusable(skip) and skip.shape == y.shape and skip.dtype == y.dtype:
            from . import _lib
            L = _lib.lib()
            dt = 1 if y.dtype == torch.bfloat16 else 0       # RCNN_BF16 / RCNN_F32
            w1, w2 = getattr(self, f"se{k}_w1"), getattr(self, f"se{k}_w2")
            yb = getattr(self, f"bf{c2[0]}")
            sb = getattr(self, f"bf{ds[0]}") if ds is not None else None
            with torch.cuda.device(y.device):
                key = (B, C, y.device)
                ws = self._ws.get(key)
                if ws is None:
                    ws = self._ws[key] = torch.zeros(int(L.rcnn_se_gate_workspace_bytes(B, C)), dtype=torch.uint8, device=y.device)
                gate = torch.empty((B, C), dtype=torch.float32, device=y.device)
                out = torch.empty_like(y)                    # (channels_last, like y)
                s = _lib.stream_ptr()
                _lib.check(L.rcnn_se_gate(y.data_ptr(), dt, B, H * W, C, w1.data_ptr(), w2.data_ptr(), w1.shape[0], yb.data_ptr(),
                                          gate.data_ptr(), ws.data_ptr(), s), "rcnn_se_gate")
                _lib.check(L.rcnn_se_apply(y.data_ptr(), skip.data_ptr(), gate.data_ptr(), yb.data_ptr(),
                                           sb.data_ptr() if sb is not None else None, dt, B, H * W, C, out.data_ptr(), s),
                           "rcnn_se_apply")
            return out
        y = y + getattr(self, f"b{c2[0]}").view(1, -1, 1, 1)
        if ds is not None:
            skip = skip + getattr(self, f"b{ds[0]}").view(1, -1, 1, 1)
        y = y * fc(y.mean(dim=(2, 3)))[:, :, None, None]
        return torch.relu_(y + skip)

    @torch.no_grad()
    def forward(self, x):
        x = x.to(self.dtype)
        if self.in_pad > x.shape[1]:                         # zero channels up to 8: cuDNN's tensor-core convolution path
            xp = torch.empty((x.shape[0], self.in_pad, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                             memory_format=torch.channels_last).zero_()
            xp[:, :x.shape[1]] = x
            x = xp
        else:
            x = x.contiguous(memory_format=torch.channels_last)
        for kind, spec in self.stem:
            x = self._conv(x, spec, True) if kind == "cr" else self._pool(spec, x)
        for k, ((c1, c2, ds), fc) in enumerate(zip(self.blocks, self.se)):
            y = self._conv(self._conv(x, c1, True), c2, False, bias=False)
            x = self._se_tail(k, x, c2, ds, y, fc)
        for kind, spec in self.tail:
            x = self._conv(x, spec, True)
        return x


class RCNN(nn.Module):
    """Drop-in for model.model.RCNN (model/model.py:166-227) with a CTC head.

    Same constructor and ``forward(x, text=None, is_train=True, batch_max_length=25)`` signature;
    ``encode`` is the reference's (CNN -> mean over height -> [B,W',512] -> enc_rnn -> dropout,
    model/model.py:215-221) with ``enc_rnn`` running on the sm_100a kernels.  ``forward`` returns
    CTC logits [B, T, num_classes + 1] (class 0 = blank, class k = itos[k-1]); ``text`` and
    ``batch_max_length`` belong to the reference's attention decoder and are accepted but unused.
    """

    def __init__(self, num_classes, hidden_size=256, sos_id: int = 1, eos_id: int = 2, pad_id: int = 0,
                 blank_id=3, enc_dropout_p: float = 0.1, dropblock_p: float = 0.0, dropblock_block_size: int = 5,
                 decoder: str = "ctc"):
        super().__init__()
        if decoder not in ("ctc", "attention"):
            raise ValueError(f"decoder must be 'ctc' or 'attention', got {decoder!r}")
        self.decoder = decoder
        self.num_classes = num_classes
        self.hidden_size = hidden_size
        self.sos_id, self.eos_id, self.pad_id, self.blank_id = sos_id, eos_id, pad_id, blank_id
        self.ctc_blank = 0
        self.num_ctc_classes = num_classes + 1
        self.cnn = SEResNet31(3, 512, dropblock_p=dropblock_p, dropblock_block_size=dropblock_block_size)
        # CTC: the head consumes bf16 (it would cast anyway); attention: fp32 like the reference's encoder output
        self.enc_rnn = make_enc_rnn(self.cnn.out_channels, hidden_size,
                                    out_dtype=torch.bfloat16 if decoder == "ctc" else torch.float32)
        self.enc_dropout = nn.Dropout(enc_dropout_p)
        # decoder="attention": the reference's own decoder (model/model.py:204-214); forward() then returns what the
        # reference's forward returns, attn.* of a checkpoint is loaded too, and there is NO ctc_head, so that the
        # state dict has exactly the reference's keys (its strict=True loaders accept checkpoints saved from here)
        self.attn = None
        self.ctc_head = CTCHead(hidden_size, self.num_ctc_classes) if decoder == "ctc" else None
        if decoder == "attention":
            from .attention import Attention
            self.attn = Attention(input_size=hidden_size, hidden_size=hidden_size, num_classes=num_classes, sos_id=sos_id,
                                  eos_id=eos_id, pad_id=pad_id, blank_id=blank_id, dropout_p=0.1, sampling_prob=0.0)

    def encode_features(self, feats: torch.Tensor) -> torch.Tensor:
        """[B, T, 512] feature columns -> [B, T, H] fp32 (the hot path without the backbone)."""
        return self._encode_features(feats).float()

    def _encode_features(self, feats: torch.Tensor) -> torch.Tensor:
        # bf16 when the CTC head follows (its GEMM operand format), fp32 for the attention decoder
        return self.enc_dropout(self.enc_rnn(feats))

    def fold_backbone(self, dtype: torch.dtype = torch.bfloat16):
        """Inference: build the folded copy of the backbone (``FoldedBackbone``) from the current weights, on their device;
        ``forward`` / ``encode`` use it whenever the model is in eval() mode.  Call again after loading other weights;
        ``fold_backbone(None)`` drops it.  The copy lives outside the module tree: the state dict keeps the reference's keys."""
        folded = None
        if dtype is not None:
            folded = FoldedBackbone(self.cnn, dtype).to(next(self.cnn.parameters()).device).eval()
        object.__setattr__(self, "_folded", folded)
        return self

    def _features(self, x: torch.Tensor) -> torch.Tensor:
        folded = getattr(self, "_folded", None)
        if folded is not None and not self.training:
            return folded(x).float().mean(dim=2).permute(0, 2, 1)
        return self.cnn(x).mean(dim=2).permute(0, 2, 1)   # AdaptiveAvgPool2d((1, None)) + squeeze(2): [B, W', C]

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.encode_features(self._features(x))

    def forward(self, x, text=None, is_train=True, batch_max_length=25):
        enc = self._encode_features(self._features(x))
        if self.attn is not None:
            return self.attn(enc, text=text, is_train=is_train, batch_max_length=batch_max_length)
        return self.ctc_head(enc)

    def load_reference_state_dict(self, state_dict, strict_encoder: bool = True):
        """Load ``cnn.*`` and ``enc_rnn.*`` (and ``attn.*`` when built with decoder="attention") from a
        reference checkpoint; the CTC head keeps its own weights.  Returns the keys that were not used."""
        own = self.state_dict()
        prefixes = ("cnn.", "enc_rnn.", "attn.") if self.attn is not None else ("cnn.", "enc_rnn.")
        picked = {k: v for k, v in state_dict.items() if k.startswith(prefixes)}
        if strict_encoder:
            missing = [k for k in own if k.startswith(prefixes) and k not in picked]
            if missing:
                raise KeyError(f"reference state dict lacks encoder keys: {missing[:5]}...")
        own.update(picked)
        self.load_state_dict(own, strict=True)
        return [k for k in state_dict if k not in picked]
