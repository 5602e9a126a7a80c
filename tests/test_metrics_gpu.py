"""GPU parity: validation metrics (K5, batched edit distance) vs the restated reference metrics
(oracle/host_ref.py: training/metrics.py:5-32).  Integer work: bit-exact, the rates compare with ==."""
import glob
import json
import os
import random

import numpy as np
import pytest
import torch

import rcnn_ocr_b200 as R
from conftest import GOLDEN, golden
from oracle import host_ref

pytestmark = pytest.mark.gpu

ALPHABET = ["<PAD>", "<SOS>", "<EOS>", " "] + list("abcdefghijklmnopqrstuvwxyzАБВГДЕЖЗ0123456789.,-")


def _strings(ids_rows, alphabet):
    return ["".join(alphabet[k - 1] for k in row if 0 < k <= len(alphabet)) for row in ids_rows]


def _device_inputs(hyps, refs, T, padded):
    N = len(hyps)
    ids = torch.full((N, T), -1, dtype=torch.int32)
    for n, h in enumerate(hyps):
        ids[n, : len(h)] = torch.tensor(h, dtype=torch.int32)
    lens = torch.tensor([len(h) for h in hyps], dtype=torch.int32)
    tl = torch.tensor([len(r) for r in refs], dtype=torch.int64)
    if padded:
        S = max(1, max((len(r) for r in refs), default=1))
        tg = torch.zeros((N, S), dtype=torch.int64)
        for n, r in enumerate(refs):
            tg[n, : len(r)] = torch.tensor(r, dtype=torch.int64)
    else:
        tg = torch.tensor([k for r in refs for k in r], dtype=torch.int64)
    return ids.cuda(), lens.cuda(), tg.cuda(), tl


def _random_pairs(N, T, seed, classes):
    rng = random.Random(seed)
    hyps, refs = [], []
    for n in range(N):
        lr = rng.choice([0, 1, 2, 5, 17, 32, T])
        ref = [rng.choice(classes) for _ in range(lr)]
        mode = rng.random()
        if mode < 0.25:
            hyp = list(ref)                                   # exact match
        elif mode < 0.75:                                     # a few edits of the reference
            hyp = list(ref)
            for _ in range(rng.randint(1, 4)):
                op = rng.random()
                pos = rng.randint(0, len(hyp))
                if op < 0.34 and hyp:
                    hyp.pop(min(pos, len(hyp) - 1))
                elif op < 0.67:
                    hyp.insert(pos, rng.choice(classes))
                elif hyp:
                    hyp[min(pos, len(hyp) - 1)] = rng.choice(classes)
            hyp = hyp[:T]
        else:
            hyp = [rng.choice(classes) for _ in range(rng.randint(0, T))]
        hyps.append(hyp[:T])
        refs.append(ref)
    return hyps, refs


@pytest.mark.parametrize("N,T,padded,special", [(1, 4, True, False), (64, 32, True, False), (257, 64, False, False),
                                                (100, 64, True, True), (33, 16, False, True)])
def test_cer_wer_accuracy_match_reference_metrics(N, T, padded, special):
    # classes 5.. are single characters; class 4 is the space; 1..3 are the multi-character special tokens
    classes = list(range(1 if special else 4, len(ALPHABET) + 1))
    classes += [4] * 6                                        # plenty of spaces -> several words per line
    hyps, refs = _random_pairs(N, T, seed=N * 7 + T, classes=classes)
    ids, lens, tg, tl = _device_inputs(hyps, refs, T, padded)
    table = R.CharsetTable(ALPHABET, ids.device)
    hs, rs = _strings(hyps, ALPHABET), _strings(refs, ALPHABET)
    dist, nref, nhyp = R.edit_stats(ids, lens, tg, tl, table)
    assert dist.cpu().tolist() == [host_ref.levenshtein(r, h) for r, h in zip(rs, hs)]
    assert nref.cpu().tolist() == [len(r) for r in rs] and nhyp.cpu().tolist() == [len(h) for h in hs]
    assert R.character_error_rates(ids, lens, tg, tl, table) == [host_ref.character_error_rate(r, h) for r, h in zip(rs, hs)]
    assert R.word_error_rates(ids, lens, tg, tl, table) == [host_ref.word_error_rate(r, h) for r, h in zip(rs, hs)]
    m = R.validation_metrics(ids, lens, tg, tl, table)
    assert m["accuracy"] == host_ref.compute_accuracy(rs, hs)
    assert m["cer"] == sum(host_ref.character_error_rate(r, h) for r, h in zip(rs, hs)) / max(1, N)
    assert m["wer"] == sum(host_ref.word_error_rate(r, h) for r, h in zip(rs, hs)) / max(1, N)


def test_metrics_from_the_reference_decoders_golden_outputs():
    """decode (K4) -> metrics (K5) entirely on the device equals the reference's strings -> metrics path:
    logits and decoded strings come from the reference's own ctc_greedy_decoder (tests/golden)."""
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "decode_*.npz")))
    checked = 0
    for name in names:
        d = golden(name)
        alphabet = json.loads(str(d["alphabet"]))
        blank = int(d["blank"]) if "blank" in d.files else 0
        if blank != 0:
            continue                                          # the metric table maps class k -> alphabet[k-1] with blank 0
        logits = torch.from_numpy(d["logits"]).cuda()
        if logits.dim() == 3 and logits.shape[0] < logits.shape[1]:
            logits = logits.permute(1, 0, 2)                  # the reference's [T,B,C] heuristic (utils.py:132-133)
        ref_texts, ref_seqs = json.loads(str(d["texts"])), json.loads(str(d["seqs"]))
        ids, lens = R.ctc_greedy_ids(logits, blank=0)
        # references: the reference decoder's own sequences with a deterministic corruption
        rng = random.Random(len(name))
        tgt = [[k for k in s if rng.random() > 0.2] for s in ref_seqs]
        tl = torch.tensor([len(t) for t in tgt], dtype=torch.int64)
        tg = torch.tensor([k for t in tgt for k in t] or [0], dtype=torch.int64).cuda()
        table = R.CharsetTable(alphabet, ids.device)
        want_refs = _strings(tgt, alphabet)
        got = R.character_error_rates(ids, lens, tg, tl, table)
        assert got == [host_ref.character_error_rate(r, h) for r, h in zip(want_refs, ref_texts)], name
        m = R.validation_metrics(ids, lens, tg, tl, table)
        assert m["accuracy"] == host_ref.compute_accuracy(want_refs, ref_texts), name
        checked += 1
    assert checked >= 4


def test_too_long_sequence_raises_and_cpu_tensors_rejected():
    table = R.CharsetTable(["<PAD>"] * 4, torch.device("cuda"))
    ids = torch.ones((1, 80), dtype=torch.int32, device="cuda")          # 80 x 5 characters > 320
    lens = torch.tensor([80], dtype=torch.int32, device="cuda")
    tg = torch.ones((1, 2), dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError):
        R.character_error_rates(ids, lens, tg, torch.tensor([2]), table)
    with pytest.raises(RuntimeError):
        R.edit_stats(ids.cpu(), lens.cpu(), tg, torch.tensor([2]), table)
