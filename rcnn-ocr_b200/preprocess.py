"""The reference's host->device input step on the device (kernel K7).

``LinePreprocessor`` replaces, for a whole batch at once, what the reference does per image on the host:
``ResizeAndPadA`` (data/transforms.py:62-120) + ``A.Normalize(0.5, 0.5)`` + ``ToTensorV2`` (:173-182, the validation
transform ``get_val_transform``) + the per-image ``.to(device)`` and ``torch.stack`` of inference.py:93-124,159-164.
Accepted inputs are the reference's (inference.py:104-119): a file path (read with ``cv2.imread``; ``FileNotFoundError``
/ ``ValueError`` as there), a ``PIL.Image`` (converted to RGB), or a uint8 ``numpy`` array [H,W] / [H,W,3] RGB /
[H,W,4] RGBA.  Decoding files stays on the host (it is I/O); the decoded pixels travel as one pinned buffer and one copy.

``pack_ctc_targets`` / ``pack_attention_targets`` are the vectorised target packers: the label strings of a batch ->
the CTC loss's concatenated class ids + lengths (class k = ``itos[k-1]``, charset.py), and the reference's
``pack_attention_targets`` (data/transforms.py:123-157; used by ``make_collate_attn``, data/dataset.py:147-156)."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

_ALIGN = {"left": 0, "top": 0, "center": 1, "right": 2, "bottom": 2}


def _load(image):
    """-> (uint8 array [H,W] / [H,W,3] / [H,W,4], bgr flag); the accepted types and errors of inference.py:104-119."""
    if isinstance(image, (str, os.PathLike)):
        path = os.fspath(image)
        if not os.path.exists(path):
            raise FileNotFoundError(f"Image file not found: {path}")
        import cv2
        img = cv2.imread(path)                      # BGR; the kernel swaps the channels while it reads
        if img is None:
            raise ValueError(f"Cannot read image: {path}")
        return img, 1
    if isinstance(image, np.ndarray):
        img = image
    else:
        try:
            from PIL import Image
        except ImportError:                          # pragma: no cover
            Image = None
        if Image is not None and isinstance(image, Image.Image):
            img = np.asarray(image.convert("RGB"))
        else:
            raise ValueError(f"Unsupported image type: {type(image)}")
    if img.dtype != np.uint8:
        raise ValueError(f"images must be uint8 (got {img.dtype}): the reference's Normalize assumes max_pixel_value 255")
    if img.ndim == 2 or (img.ndim == 3 and img.shape[2] in (1, 3, 4)):
        return img, 0
    raise ValueError(f"Unsupported image shape: {img.shape}")


class LinePreprocessor:
    """callable(images) -> float32 (or bf16) tensor [N, 3, img_h, img_w] on ``device``, normalised to [-1, 1]."""

    def __init__(self, img_h: int = 32, img_w: int = 256, device="cuda", align_h: str = "left", align_v: str = "center",
                 dtype: torch.dtype = torch.float32):
        self.img_h, self.img_w = int(img_h), int(img_w)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("rcnn-ocr_b200 has no CPU path; pass a CUDA device")
        self.align = (_ALIGN[align_h], _ALIGN[align_v])
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be float32 or bfloat16")
        self.dtype = dtype
        self._pinned = None        # grows as needed; reused across calls (a call returns after its copy was enqueued and
        self._pinned_desc = None   # the next call waits for it: see _copied)
        self._copied = None

    def pack(self, images: Sequence):
        """Host half: decode / view every image and lay them out in one pinned buffer.  Returns (pixels u8 pinned,
        desc int64 [N,6] pinned, N)."""
        loaded = [_load(im) for im in images]
        sizes = [int(a.shape[0]) * int(a.shape[1]) * (1 if a.ndim == 2 else int(a.shape[2])) for a, _ in loaded]
        offs = np.zeros(len(loaded) + 1, dtype=np.int64)
        np.cumsum(np.asarray(sizes, dtype=np.int64) + 15 & ~15, out=offs[1:])     # 16-byte aligned starts
        total = int(offs[-1]) if loaded else 0
        if self._copied is not None:
            self._copied.synchronize()               # the previous call's H2D copy has read the pinned buffers
        if self._pinned is None or self._pinned.numel() < max(total, 1):
            self._pinned = torch.empty(max(total, 1) * 5 // 4 + 16, dtype=torch.uint8).pin_memory()
        if self._pinned_desc is None or self._pinned_desc.shape[0] < max(len(loaded), 1):
            self._pinned_desc = torch.empty((max(len(loaded), 1) * 5 // 4 + 1, 6), dtype=torch.int64).pin_memory()
        buf = self._pinned.numpy()
        desc = self._pinned_desc.numpy()
        for i, (a, bgr) in enumerate(loaded):
            h, w = int(a.shape[0]), int(a.shape[1])
            ch = 1 if a.ndim == 2 else int(a.shape[2])
            if h == 0 or w == 0:
                raise ValueError("empty image")
            buf[offs[i]:offs[i] + sizes[i]] = np.ascontiguousarray(a).reshape(-1)
            desc[i] = (offs[i], h, w, w * ch, ch, bgr)
        return self._pinned[:max(total, 1)], self._pinned_desc[:max(len(loaded), 1)], len(loaded)

    def __call__(self, images: Sequence) -> torch.Tensor:
        pixels, desc, n = self.pack(list(images))
        with torch.cuda.device(self.device):
            out = torch.empty((n, 3, self.img_h, self.img_w), dtype=self.dtype, device=self.device)
            if n == 0:
                return out
            dpix = pixels.to(self.device, non_blocking=True)
            ddesc = desc.to(self.device, non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record()
            done = 0
            while done < n:                          # (grid.y limit)
                m = min(n - done, 65535)
                rc = _lib.lib().rcnn_preprocess_lines(dpix.data_ptr(), ddesc[done:].data_ptr(), m, self.img_h, self.img_w,
                                                      self.align[0], self.align[1], out[done:].data_ptr(),
                                                      0 if self.dtype == torch.float32 else 1, _lib.stream_ptr())
                _lib.check(rc, "rcnn_preprocess_lines")
                done += m
        return out


def _label_ids(texts: Sequence[str], stoi: dict, drop_blank: bool, max_len: int | None):
    blank = stoi.get("<BLANK>") if drop_blank else None
    get = stoi.get
    rows = []
    for s in texts:
        ids = [i for i in map(get, s) if i is not None and i != blank]     # characters outside the charset are skipped
        rows.append(ids[:max_len] if max_len is not None else ids)
    return rows


def pack_ctc_targets(texts: Sequence[str], stoi: dict, max_len: int | None = None, drop_blank: bool = True):
    """Label strings -> (targets int64 1-D, concatenated CTC class ids; target_lengths int64 [N]).  CTC class k is
    ``itos[k-1]`` (class 0 = blank, charset.py), characters that are not in the charset are skipped and ``<BLANK>`` is
    dropped exactly as ``pack_attention_targets`` does (data/transforms.py:139-147); ``max_len`` truncates likewise."""
    rows = _label_ids(texts, stoi, drop_blank, max_len)
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    flat = np.fromiter((i + 1 for r in rows for i in r), dtype=np.int64, count=int(lens.sum()))
    return torch.from_numpy(flat), torch.from_numpy(lens)


def pack_attention_targets(texts: Sequence[str], stoi: dict, max_len: int, drop_blank: bool = True):
    """data/transforms.py:123-157, same returns (text_in [B, max_len+1] with <SOS> first, target_y [B, max_len+1] with
    <EOS> after the characters, lengths = characters + 1), built with two array writes instead of per-sample tensors."""
    PAD, SOS, EOS = stoi["<PAD>"], stoi["<SOS>"], stoi["<EOS>"]
    rows = _label_ids(texts, stoi, drop_blank, max_len)
    Bn, Tn = len(rows), max_len + 1
    text_in = np.full((Bn, Tn), PAD, dtype=np.int64)
    target_y = np.full((Bn, Tn), PAD, dtype=np.int64)
    if Bn:
        text_in[:, 0] = SOS
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=Bn)
    if lens.sum() > 0:
        flat = np.fromiter((i for r in rows for i in r), dtype=np.int64, count=int(lens.sum()))
        bidx = np.repeat(np.arange(Bn), lens)
        pos = np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens)
        text_in[bidx, pos + 1] = flat
        target_y[bidx, pos] = flat
    if Bn:
        target_y[np.arange(Bn), lens] = EOS
    return torch.from_numpy(text_in), torch.from_numpy(target_y), torch.from_numpy(lens + 1)
