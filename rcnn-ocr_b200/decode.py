"""Greedy CTC decode on the device (kernel K4) behind the reference's decode surface.

Mirrors training/utils.py:122-162: ``ctc_greedy_decoder(logits, alphabet, blank=0)`` returns
``(texts, seqs)`` and ``decode(ctc_out, alphabet, method="greedy")`` unwraps tuples and
rejects unknown methods with ValueError.  The reference's layout heuristic (permute when
``shape[0] < shape[1]``, utils.py:132-133) is kept as the default; ``batch_first`` makes the
layout explicit.  Only ids[B,T] int32 and len[B] cross back to the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def ctc_greedy_ids(logits: torch.Tensor, blank: int = 0, batch_first: bool = True,
                   return_confidence: bool = False):
    """argmax + collapse + strip blank.  logits [B,T,C] (or [T,B,C] with batch_first=False),
    float32 or bfloat16, on a CUDA device (host tensors are copied over first).
    Returns device tensors (ids [B,T] int32 left-packed / -1 padded, lens [B] int32[, conf [B]])."""
    if logits.dim() != 3:
        raise ValueError(f"expected a 3-D logits tensor, got shape {tuple(logits.shape)}")
    if not logits.is_cuda:
        src = logits if logits.is_pinned() else logits.contiguous().pin_memory()
        logits = src.to("cuda", non_blocking=True)
    if logits.dtype not in _DTYPES:
        logits = logits.float()
    if not batch_first:
        logits = logits.permute(1, 0, 2)
    if logits.stride(2) != 1 and logits.shape[2] > 1:
        logits = logits.contiguous()
    B, T, C = logits.shape
    if C == 0:
        raise ValueError("logits have zero classes")
    with torch.cuda.device(logits.device):
        ids = torch.empty((B, T), dtype=torch.int32, device=logits.device)
        lens = torch.empty((B,), dtype=torch.int32, device=logits.device)
        conf = torch.empty((B,), dtype=torch.float32, device=logits.device) if return_confidence else None
        rc = _lib.lib().rcnn_ctc_greedy(
            logits.data_ptr(), _DTYPES[logits.dtype], B, T, C, logits.stride(0), logits.stride(1),
            int(blank), ids.data_ptr(), lens.data_ptr(), conf.data_ptr() if conf is not None else None,
            _lib.stream_ptr())
        _lib.check(rc, "rcnn_ctc_greedy")
    return (ids, lens, conf) if return_confidence else (ids, lens)


def _to_host(ids: torch.Tensor, lens: torch.Tensor):
    """One packed device->host copy of ids and lens."""
    B, T = ids.shape
    packed = torch.empty((B, T + 1), dtype=torch.int32, device=ids.device)
    packed[:, :T] = ids
    packed[:, T] = lens
    host = torch.empty((B, T + 1), dtype=torch.int32, pin_memory=True)
    host.copy_(packed, non_blocking=True)
    torch.cuda.current_stream(ids.device).synchronize()
    arr = host.numpy()
    return arr[:, :T], arr[:, T]


def ids_to_text(ids: torch.Tensor, lens: torch.Tensor, alphabet):
    """Device ids [B,T] / lens [B] (as returned by ctc_greedy_ids) -> (texts, seqs): one packed D2H
    copy, then ``alphabet[p - 1]`` per emitted class as at training/utils.py:146."""
    ids_h, lens_h = _to_host(ids, lens)
    table = np.asarray(list(alphabet), dtype=object)
    seqs, texts = [], []
    for b in range(ids_h.shape[0]):
        row = ids_h[b, : lens_h[b]]
        seqs.append(row.tolist())
        texts.append("".join(table[row - 1]) if len(row) else "")
    return texts, seqs


def ctc_greedy_decoder(logits: torch.Tensor, alphabet, blank: int = 0, batch_first=None):
    """Drop-in for training/utils.py:122-150.  Returns (texts, seqs)."""
    if batch_first is None:  # the reference's heuristic, utils.py:132-133
        batch_first = not (logits.dim() == 3 and logits.shape[0] < logits.shape[1])
    ids, lens = ctc_greedy_ids(logits, blank=blank, batch_first=batch_first)
    return ids_to_text(ids, lens, alphabet)


def decode(ctc_out, alphabet, method: str = "greedy"):
    """Drop-in for training/utils.py:153-162.  The reference applies log_softmax before the
    argmax; it is argmax-invariant, so the kernel reads the raw scores."""
    if isinstance(ctc_out, tuple):
        ctc_out = ctc_out[0]
    if method == "greedy":
        return ctc_greedy_decoder(ctc_out, alphabet)
    raise ValueError(f"Unsupported decode method: {method}")
