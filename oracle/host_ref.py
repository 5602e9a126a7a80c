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


# ---- input step (restated; cv2 and numpy are the reference's own dependencies and are present) -------------------
#   resize_and_pad        data/transforms.py:91-120  (ResizeAndPadA.apply)
#   normalize_chw         data/transforms.py:179 + ToTensorV2: albumentations' normalize is float32
#                         (img - mean * 255) * reciprocal(std * 255), then HWC -> CHW
#   pack_attention_targets data/transforms.py:123-157
# Pinned by tests/golden/preproc_*.npz and pack_attn_*.npz, which tests/make_golden.py wrote by calling the reference's
# own ResizeAndPadA.apply / pack_attention_targets (albumentations stubbed for the import).

def resize_and_pad(img, img_h: int, img_w: int, align_h: str = "left", align_v: str = "center"):
    import cv2
    import numpy as np
    if img.ndim == 2:
        img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
    elif img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_RGBA2RGB)
    elif img.shape[2] == 1:
        img = cv2.cvtColor(img[:, :, 0], cv2.COLOR_GRAY2RGB)
    h, w = img.shape[:2]
    scale = min(img_h / max(h, 1), img_w / max(w, 1))
    new_w = max(1, int(round(w * scale)))
    new_h = max(1, int(round(h * scale)))
    interp = cv2.INTER_AREA if (new_h < h or new_w < w) else cv2.INTER_LINEAR
    resized = cv2.resize(img, (new_w, new_h), interpolation=interp)
    canvas = np.full((img_h, img_w, 3), 255, dtype=img.dtype)
    x0 = 0 if align_h == "left" else (img_w - new_w if align_h == "right" else (img_w - new_w) // 2)
    y0 = 0 if align_v == "top" else (img_h - new_h if align_v == "bottom" else (img_h - new_h) // 2)
    x0 = max(0, min(x0, img_w - new_w))
    y0 = max(0, min(y0, img_h - new_h))
    canvas[y0:y0 + new_h, x0:x0 + new_w] = resized
    return canvas


def normalize_chw(canvas):
    import numpy as np
    out = (canvas.astype(np.float32) - np.float32(127.5)) * np.reciprocal(np.float32(127.5))
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def pack_attention_targets(texts, stoi, max_len, drop_blank=True):
    import numpy as np
    PAD, SOS, EOS = stoi["<PAD>"], stoi["<SOS>"], stoi["<EOS>"]
    BLANK = stoi.get("<BLANK>")
    n, t = len(texts), max_len + 1
    text_in = np.full((n, t), PAD, dtype=np.int64)
    text_in[:, 0] = SOS
    target_y = np.full((n, t), PAD, dtype=np.int64)
    lengths = np.zeros(n, dtype=np.int64)
    for i, s in enumerate(texts):
        ids = []
        for ch in s:
            if ch not in stoi:
                continue
            k = stoi[ch]
            if drop_blank and BLANK is not None and k == BLANK:
                continue
            ids.append(k)
        m = min(len(ids), max_len)
        text_in[i, 1:1 + m] = ids[:m]
        target_y[i, :m] = ids[:m]
        target_y[i, m] = EOS
        lengths[i] = m + 1
    return text_in, target_y, lengths
