"""Pure-Python restatements of the small host-side pieces next to the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference modules that hold
these cannot be imported in this image (albumentations / Levenshtein / jiwer are
missing), so they are restated from the source:

  load_charset            data/transforms.py:39-59
  decode_tokens           data/transforms.py:196-206
  character_error_rate    training/metrics.py:5-13   (own Levenshtein DP)
  compute_accuracy        training/metrics.py:23-32
  word_error_rate         training/metrics.py:16-21  (jiwer default word split, restated)
"""
from __future__ import annotations


def load_charset(path: str):
    itos = []
    with open(path, "r", encoding="utf-8") as fh:
        for raw in fh:
            tok = raw.rstrip("\n")
            if tok != "":
                itos.append(tok)
    return itos, {s: i for i, s in enumerate(itos)}


def decode_tokens(ids, itos, pad_id, eos_id, blank_id=None):
    chars = []
    for t in ids:
        t = int(t)
        if t == eos_id:
            break
        if t == pad_id or (blank_id is not None and t == blank_id):
            continue
        chars.append(itos[t])
    return "".join(chars)


def levenshtein(a, b) -> int:
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def character_error_rate(reference: str, hypothesis: str) -> float:
    if len(reference) == 0:
        return float("inf") if len(hypothesis) > 0 else 0.0
    return levenshtein(reference, hypothesis) / len(reference)


def compute_accuracy(references, hypotheses) -> float:
    if len(references) == 0:
        return 0.0
    return sum(1 for r, h in zip(references, hypotheses) if r == h) / len(references)


def word_error_rate(reference: str, hypothesis: str) -> float:
    """training/metrics.py:16-21 calls jiwer.wer(reference, hypothesis) (jiwer is not installed in this
    image: third-party, pinned by requirements.txt as ``jiwer``).  Its default transform collapses runs
    of whitespace, strips the ends and splits on " "; WER = word-level Levenshtein / number of reference
    words.  jiwer raises ValueError for an empty reference; the restatement (and the device kernel) use
    the CER convention there: inf for a non-empty hypothesis, 0 for two empty strings."""
    ref_words = [w for w in reference.split(" ") if w != ""]
    hyp_words = [w for w in hypothesis.split(" ") if w != ""]
    if len(ref_words) == 0:
        return float("inf") if len(hyp_words) > 0 else 0.0
    return levenshtein(ref_words, hyp_words) / len(ref_words)
