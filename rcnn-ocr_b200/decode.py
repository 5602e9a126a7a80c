"""Greedy CTC decode on the device (kernel K4) behind the reference's decode surface.

Mirrors training/utils.py:122-162: ``ctc_greedy_decoder(logits, alphabet, blank=0)`` returns
``(texts, seqs)`` and ``decode(ctc_out, alphabet, method="greedy")`` unwraps tuples and
rejects unknown methods with ValueError.  The reference's layout heuristic (permute when
``shape[0] < shape[1]``, utils.py:132-133) is kept as the default; ``batch_first`` makes the
layout explicit.  Only ids[B,T] int32 and len[B] cross back to the host.
"""
from __future__ import annotations

import functools

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


_PINNED = {}     # (B, T, device) -> (pinned ids buffer, pinned lens buffer); a handful of batch shapes per process


def _to_host(ids: torch.Tensor, lens: torch.Tensor):
    """Device->host copy of ids and lens through cached pinned buffers (two async copies, one sync);
    returns numpy arrays the caller owns."""
    B, T = ids.shape
    key = (B, T, ids.device.index)
    if key not in _PINNED:
        if len(_PINNED) >= 16:
            _PINNED.clear()
        _PINNED[key] = (torch.empty((B, T), dtype=torch.int32, pin_memory=True),
                        torch.empty((B,), dtype=torch.int32, pin_memory=True))
    h_ids, h_lens = _PINNED[key]
    h_ids.copy_(ids if ids.is_contiguous() else ids.contiguous(), non_blocking=True)
    h_lens.copy_(lens, non_blocking=True)
    torch.cuda.current_stream(ids.device).synchronize()
    return h_ids.numpy().copy(), h_lens.numpy().copy()


@functools.lru_cache(maxsize=8)
def _alphabet_tables(alphabet: tuple):
    """(code points as uint32 if every class is a single character else None, object table)."""
    single = all(isinstance(a, str) and len(a) == 1 for a in alphabet)
    cps = np.array([ord(a) for a in alphabet], dtype=np.uint32) if single and alphabet else None
    return cps, np.asarray(list(alphabet), dtype=object)


def ids_to_text_host(ids_h: np.ndarray, lens_h: np.ndarray, alphabet):
    """Host half of the decode: ids [B,T] int32 (left-packed) / lens [B] -> (texts, seqs) with
    ``alphabet[p - 1]`` per emitted class as at training/utils.py:146.  The whole batch is gathered and mapped at
    once (single-character alphabets: one code-point gather + one utf-32 decode), then cut per line."""
    B, T = ids_h.shape
    lens_l = lens_h.tolist()
    if B == 0 or T == 0:
        return ["" for _ in lens_l], [[] for _ in lens_l]
    mask = np.arange(T, dtype=np.int64)[None, :] < lens_h.astype(np.int64)[:, None]
    flat = ids_h[mask]
    # (indexing as in the reference: class 0, possible only with blank != 0, wraps to alphabet[-1]; ids beyond
    # the alphabet raise IndexError)
    fl = flat.tolist()
    cps, table = _alphabet_tables(tuple(alphabet))
    seqs, texts, o = [], [], 0
    if cps is not None:
        allc = cps[flat - 1].tobytes().decode("utf-32-le")
        for n in lens_l:
            seqs.append(fl[o:o + n])
            texts.append(allc[o:o + n])
            o += n
    else:
        toks = table[flat - 1].tolist() if flat.size else []
        for n in lens_l:
            seqs.append(fl[o:o + n])
            texts.append("".join(toks[o:o + n]))
            o += n
    return texts, seqs


def ids_to_text(ids: torch.Tensor, lens: torch.Tensor, alphabet):
    """Device ids [B,T] / lens [B] (as returned by ctc_greedy_ids) -> (texts, seqs): D2H of ids and lens,
    then the charset mapping on the host (``ids_to_text_host``)."""
    ids_h, lens_h = _to_host(ids, lens)
    return ids_to_text_host(ids_h, lens_h, alphabet)


class PendingTexts:
    """The host half of a decode, deferred: the device->host copies of ids / lens are enqueued (own pinned buffers, own
    event) when the object is made, ``result()`` waits for that event only and maps through the charset.  A serving
    loop launches batch i+1 before it asks for batch i's strings, so the copy and the Python work of one batch run
    under the kernels of the next (``ids_to_text`` = the same with an immediate ``result()``)."""

    _pool = {}

    def __init__(self, ids: torch.Tensor, lens: torch.Tensor, alphabet):
        B, T = ids.shape
        key = (B, T, ids.device.index)
        ring = PendingTexts._pool.setdefault(key, {"next": 0, "bufs": []})
        if len(ring["bufs"]) < 4:
            ring["bufs"].append((torch.empty((B, T), dtype=torch.int32, pin_memory=True),
                                 torch.empty((B,), dtype=torch.int32, pin_memory=True), torch.cuda.Event()))
        self.h_ids, self.h_lens, self.event = ring["bufs"][ring["next"] % len(ring["bufs"])]
        ring["next"] += 1
        self.alphabet = alphabet
        self.h_ids.copy_(ids if ids.is_contiguous() else ids.contiguous(), non_blocking=True)
        self.h_lens.copy_(lens, non_blocking=True)
        self.event.record(torch.cuda.current_stream(ids.device))

    def result(self):
        self.event.synchronize()
        return ids_to_text_host(self.h_ids.numpy(), self.h_lens.numpy(), self.alphabet)


def ids_to_text_async(ids: torch.Tensor, lens: torch.Tensor, alphabet) -> PendingTexts:
    """Enqueue the D2H of a decoded batch and return a handle; ``.result()`` gives ``(texts, seqs)``.  At most four
    handles per batch shape may be outstanding (their pinned buffers rotate)."""
    return PendingTexts(ids, lens, alphabet)


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
