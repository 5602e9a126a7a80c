"""CTC loss on the device (kernel K3) behind torch's nn.CTCLoss surface.

The reference trains with ``criterion(logits2d, target1d)`` at training/train.py:289,503-505;
the north_star replaces that call with an nn.CTCLoss-compatible loss:

    CTCLoss(blank=0, reduction='mean', zero_infinity=False)(log_probs[T,N,C], targets,
                                                            input_lengths, target_lengths)

``ctc_loss_from_logits`` is the fused form the training step uses: it takes the raw logits,
applies log_softmax inside the kernel and returns the gradient w.r.t. the logits from the same
launch.  Semantics follow torch: 'mean' divides each sample's nll by max(target_len, 1) and
averages over the batch; infeasible alignments give inf (0 loss / 0 grad with zero_infinity);
frames beyond input_length get zero gradient; 2-D padded and 1-D concatenated targets.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

_REDUCTIONS = {"none": 0, "mean": 1, "sum": 2}


def _lengths(v, n: int, device, name: str):
    """Lengths as (int64 device tensor, host max or None)."""
    if isinstance(v, torch.Tensor):
        if v.dtype.is_floating_point or v.dtype == torch.bool:
            raise RuntimeError(f"{name} must be integral")
        if v.numel() != n:
            raise RuntimeError(f"{name} must have {n} elements, got {v.numel()}")
        hmax = int(v.max()) if (not v.is_cuda and n > 0) else None
        return v.reshape(n).to(device=device, dtype=torch.int64, non_blocking=True).contiguous(), hmax
    vals = [int(a) for a in v]
    if len(vals) != n:
        raise RuntimeError(f"{name} must have {n} elements, got {len(vals)}")
    return torch.tensor(vals, dtype=torch.int64, device=device), (max(vals) if vals else 0)


def sum_check(target_lengths, tg_numel: int) -> bool:
    """Concatenated targets with host-resident lengths: True when the lengths overrun the target vector
    (torch raises for this; device-resident lengths are checked by the kernel, which flags the row as bad)."""
    if isinstance(target_lengths, torch.Tensor):
        if target_lengths.is_cuda:
            return False
        total = int(target_lengths.sum())
    else:
        total = sum(int(a) for a in target_lengths)
    return total > tg_numel


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, targets, input_lengths, target_lengths, blank, reduction, zero_infinity,
                from_logits, max_target_length):
        _lib.require_cuda(x, "log_probs/logits")
        # decided BEFORE any conversion: autograd is off inside forward(), so x.float() / x.contiguous() below
        # return tensors with requires_grad == False (bf16 / fp16 log-probs under autocast, strided classes)
        need_grad = ctx.needs_input_grad[0]
        if x.dim() != 3:
            raise RuntimeError(f"expected [T,N,C] input, got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        if x.shape[2] > 1 and x.stride(2) != 1:
            x = x.contiguous()
        T, N, C = x.shape
        dev = x.device
        if not 0 <= blank < C:
            raise RuntimeError("blank must be in label range")
        il, _ = _lengths(input_lengths, N, dev, "input_lengths")
        tl, tl_max = _lengths(target_lengths, N, dev, "target_lengths")
        tg = targets.to(device=dev, dtype=torch.int64, non_blocking=True)
        if tg.dim() == 2:
            if tg.shape[0] != N:
                raise RuntimeError(f"targets must have {N} rows, got {tg.shape[0]}")
            tg = tg.contiguous()
            tgt_stride = tg.shape[1]
            bound = tg.shape[1]
            if tgt_stride == 0:          # all targets empty
                tg, tgt_stride = torch.zeros((N, 1), dtype=torch.int64, device=dev), 1
        elif tg.dim() == 1:
            tg = tg.contiguous()
            tgt_stride, bound = 0, None
            tg_numel = tg.numel()
            if tg.numel() == 0:
                tg = torch.zeros((1,), dtype=torch.int64, device=dev)
        else:
            raise RuntimeError("targets must be 1-D (concatenated) or 2-D (padded)")
        if max_target_length is None:
            if tl_max is not None:
                max_target_length = tl_max if bound is None else min(tl_max, bound)
            elif bound is not None:
                max_target_length = bound
            else:  # concatenated targets with device-resident lengths: one host sync (maximum and sum together)
                mx_sum = torch.stack([tl.max(), tl.sum()]).tolist() if N > 0 else [0, 0]
                max_target_length = int(mx_sum[0])
                if mx_sum[1] > tg_numel:
                    raise RuntimeError("sum(target_lengths) exceeds the number of concatenated targets")
        max_target_length = max(int(max_target_length), 0)
        if bound is not None:
            max_target_length = min(max_target_length, bound)   # never read past a padded row
        elif tl_max is not None and sum_check(target_lengths, tg_numel):
            raise RuntimeError("sum(target_lengths) exceeds the number of concatenated targets")
        L = _lib.lib()
        with torch.cuda.device(dev):
            nll = torch.empty((N,), dtype=torch.float32, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = None
            gst = gsn = 0
            if need_grad:
                dense = x.is_contiguous() or x.permute(1, 0, 2).is_contiguous()
                grad = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=dev) if dense \
                    else torch.empty_like(x, memory_format=torch.contiguous_format)
                gst, gsn = grad.stride(0), grad.stride(1)
            wsb = L.rcnn_ctc_workspace_bytes(T, N, C, max_target_length)
            ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
            rc = L.rcnn_ctc_loss(
                x.data_ptr(), int(bool(from_logits)), T, N, C, x.stride(0), x.stride(1),
                tg.data_ptr(), tgt_stride, il.data_ptr(), tl.data_ptr(), max_target_length,
                int(blank), _REDUCTIONS[reduction], int(bool(zero_infinity)),
                nll.data_ptr(), loss.data_ptr(), grad.data_ptr() if grad is not None else None,
                gst, gsn, ws.data_ptr(), wsb, _lib.stream_ptr())
            _lib.check(rc, "rcnn_ctc_loss")
        ctx.grad = grad
        ctx.per_sample = reduction == "none"
        ctx.consumed = False
        aux = nll.clone() if reduction == "none" else nll
        ctx.mark_non_differentiable(aux)
        return (nll if reduction == "none" else loss), aux

    @staticmethod
    def backward(ctx, grad_out, _grad_nll):
        g = ctx.grad
        if g is None:
            return (None,) * 9
        if ctx.consumed:
            raise RuntimeError("the fused CTC gradient buffer was already consumed by an earlier "
                               "backward pass (retain_graph is not supported for this op)")
        ctx.consumed = True
        T, N, C = g.shape
        scale = grad_out.to(device=g.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(g.device):
            rc = _lib.lib().rcnn_ctc_scale_grad(g.data_ptr(), T, N, C, g.stride(0), g.stride(1),
                                                scale.data_ptr(), int(ctx.per_sample), _lib.stream_ptr())
            _lib.check(rc, "rcnn_ctc_scale_grad")
        return (g,) + (None,) * 8


def _prepare(x):
    unbatched = x.dim() == 2
    return (x.unsqueeze(1), True) if unbatched else (x, False)


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank: int = 0,
             reduction: str = "mean", zero_infinity: bool = False, max_target_length=None):
    """F.ctc_loss-compatible: ``log_probs`` [T,N,C] (or [T,C]) are log-probabilities."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    x, unb = _prepare(log_probs)
    if unb:
        targets = targets.unsqueeze(0) if targets.dim() == 1 else targets
        input_lengths = torch.as_tensor(input_lengths).reshape(1)
        target_lengths = torch.as_tensor(target_lengths).reshape(1)
    out, _ = _CTCFunction.apply(x, targets, input_lengths, target_lengths, int(blank), reduction,
                                bool(zero_infinity), False, max_target_length)
    return out[0] if (unb and reduction == "none") else out


def ctc_loss_from_logits(logits, targets, input_lengths, target_lengths, blank: int = 0,
                         reduction: str = "mean", zero_infinity: bool = False, max_target_length=None,
                         return_nll: bool = False):
    """Fused log_softmax + CTC.  ``logits`` [T,N,C] in any dense layout with a unit class stride
    (so ``head_out.permute(1,0,2)`` of a batch-first [N,T,C] tensor costs nothing)."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    out, nll = _CTCFunction.apply(logits, targets, input_lengths, target_lengths, int(blank), reduction,
                                  bool(zero_infinity), True, max_target_length)
    return (out, nll) if return_nll else out


class CTCLoss(nn.Module):
    """Drop-in for torch.nn.CTCLoss(blank, reduction, zero_infinity)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _REDUCTIONS:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank,
                        self.reduction, self.zero_infinity)
