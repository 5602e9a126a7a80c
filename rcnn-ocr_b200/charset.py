"""Charset handling next to the hot path (host side, plain Python).

Mirrors data/transforms.py:39-59 (load_charset) and :196-206 (decode_tokens) of the
reference.  ``ctc_alphabet`` fixes the class <-> character map of the CTC head (SURVEY.md
section 8 a8): class 0 is the CTC blank and class k is ``itos[k-1]`` -- exactly the
``alphabet[p - 1]`` lookup of training/utils.py:146 with ``alphabet = itos`` -- so the head
has len(itos)+1 = 195 classes for configs/charset.txt.
"""
from __future__ import annotations


def load_charset(charset_path: str):
    """One token per line, trailing newline stripped, empty lines skipped -> (itos, stoi)."""
    itos = []
    with open(charset_path, "r", encoding="utf-8") as fh:
        for raw in fh:
            tok = raw.rstrip("\n")
            if tok == "":
                continue
            itos.append(tok)
    stoi = {tok: idx for idx, tok in enumerate(itos)}
    return itos, stoi


def decode_tokens(ids, itos, pad_id, eos_id, blank_id=None):
    """Attention-path token decode (stop at EOS, drop PAD/BLANK, no repeat collapse)."""
    pieces = []
    for tok in ids:
        tok = int(tok)
        if tok == eos_id:
            break
        if tok == pad_id or (blank_id is not None and tok == blank_id):
            continue
        pieces.append(itos[tok])
    return "".join(pieces)


def ctc_alphabet(itos):
    """(alphabet, num_classes, blank) for the CTC head over a loaded charset."""
    return list(itos), len(itos) + 1, 0


def encode_ctc_targets(texts, stoi):
    """Label strings -> (concatenated class ids, lengths) under the ctc_alphabet map."""
    flat, lens = [], []
    for s in texts:
        ids = [stoi[ch] + 1 for ch in s]
        flat.extend(ids)
        lens.append(len(ids))
    return flat, lens
