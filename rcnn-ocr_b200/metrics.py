"""Validation metrics on the device (kernel K5) behind the reference's metric surface.

Mirrors training/metrics.py:5-32 as it is used by the validation loop (training/train.py:582-598)
and evaluate_dataset.py:104-119: per-pair character error rate (Levenshtein / len(reference), inf
for an empty reference with a non-empty hypothesis, 0 for two empty strings), word error rate
(jiwer's default word split: runs of spaces collapse, leading/trailing spaces are dropped; the same
empty-reference convention as CER instead of jiwer's ValueError) and exact-match accuracy.

The hypotheses are the class ids the greedy decoder left on the device (``ctc_greedy_ids``), the
references the CTC target ids: nothing is decoded to Python strings, one small D2H copy brings the
per-pair integers back.  Distances are computed over Unicode code points through the charset table,
so they equal ``Levenshtein.distance`` on the decoded strings exactly.
"""
from __future__ import annotations

import torch

from . import _lib


class CharsetTable:
    """class id -> code points, on the device.  Class 0 is the CTC blank (no characters), class k is
    ``alphabet[k-1]`` (training/utils.py:146); a token may have several characters."""

    def __init__(self, alphabet, device):
        offs, cps = [0, 0], []
        for tok in alphabet:
            cps.extend(ord(ch) for ch in tok)
            offs.append(len(cps))
        self.num_classes = len(alphabet) + 1
        self.cp_off = torch.tensor(offs, dtype=torch.int32, device=device)
        self.cp = torch.tensor(cps if cps else [0], dtype=torch.int32, device=device)


def _ref_layout(targets: torch.Tensor, target_lengths: torch.Tensor):
    tl = target_lengths.to(device=targets.device, dtype=torch.int64).contiguous()
    if targets.dim() == 2:
        tg = targets.to(torch.int64).contiguous()
        off = torch.arange(tg.shape[0], device=tg.device, dtype=torch.int64) * tg.shape[1]
        return tg.view(-1), off, tl
    tg = targets.to(torch.int64).contiguous()
    off = torch.cumsum(tl, 0) - tl
    return tg, off, tl


def edit_stats(ids: torch.Tensor, lens: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor,
               table: CharsetTable, words: bool = False):
    """(distance, reference length, hypothesis length) per pair, int32 device tensors; lengths count
    characters (or words).  ids/lens as returned by ``ctc_greedy_ids``; targets padded [N,S] or
    concatenated 1-D class ids with ``target_lengths`` (the CTC loss's target format)."""
    _lib.require_cuda(ids, "ids")
    _lib.require_cuda(targets, "targets")
    N = ids.shape[0]
    assert ids.dtype == torch.int32 and lens.dtype == torch.int32 and ids.stride(1) == 1
    if ids.shape[1] == 0:
        ids = torch.full((N, 1), -1, dtype=torch.int32, device=ids.device)
    tg, off, tl = _ref_layout(targets, target_lengths)
    if tg.numel() == 0:
        tg = torch.zeros((1,), dtype=torch.int64, device=ids.device)
    with torch.cuda.device(ids.device):
        out = torch.empty((3, max(N, 1)), dtype=torch.int32, device=ids.device)
        rc = _lib.lib().rcnn_edit_distance(ids.data_ptr(), ids.stride(0), lens.contiguous().data_ptr(), tg.data_ptr(),
                                           off.data_ptr(), tl.data_ptr(), N, table.cp_off.data_ptr(), table.cp.data_ptr(),
                                           table.num_classes, int(words), out[0].data_ptr(), out[1].data_ptr(),
                                           out[2].data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "rcnn_edit_distance")
    return out[0, :N], out[1, :N], out[2, :N]


def _rates(dist, nref, nhyp):
    h = torch.stack([dist, nref, nhyp]).cpu().tolist()
    if any(d < 0 for d in h[0]):
        raise RuntimeError("edit distance: a sequence expands to more than 320 symbols")
    return [(d / r) if r > 0 else (float("inf") if y > 0 else 0.0) for d, r, y in zip(*h)], h[0]


def character_error_rates(ids, lens, targets, target_lengths, table: CharsetTable):
    """Per-pair CER (python floats, the values of training/metrics.py:5-13)."""
    return _rates(*edit_stats(ids, lens, targets, target_lengths, table, words=False))[0]


def word_error_rates(ids, lens, targets, target_lengths, table: CharsetTable):
    """Per-pair WER (training/metrics.py:16-21 with jiwer's default word split)."""
    return _rates(*edit_stats(ids, lens, targets, target_lengths, table, words=True))[0]


def validation_metrics(ids, lens, targets, target_lengths, table: CharsetTable):
    """The three validation numbers of training/train.py:582-584 for one set:
    accuracy = exact matches / N, CER and WER = mean of the per-pair rates (python-order sums)."""
    cers, dists = _rates(*edit_stats(ids, lens, targets, target_lengths, table, words=False))
    wers, _ = _rates(*edit_stats(ids, lens, targets, target_lengths, table, words=True))
    n = len(cers)
    acc = (sum(1 for d in dists if d == 0) / n) if n else 0.0
    return {"accuracy": acc, "cer": sum(cers) / max(1, n), "wer": sum(wers) / max(1, n)}
