"""Inference surface mirroring inference.py:12-194 of the reference (``OCRInference.predict``):
batch loop -> model -> decode -> strings.  decoder="ctc" (default): greedy CTC decode on the device (K4);
decoder="attention": the reference's own path -- greedy attention decoding on the device (K6), argmax,
``decode_tokens`` (inference.py:166-180), so existing reference checkpoints predict unchanged.

``predict`` takes what the reference's takes (inference.py:93-124): file paths, PIL images, uint8 numpy arrays
([H,W], [H,W,3] RGB, [H,W,4] RGBA) or a list of them -- resized, padded, normalised and batched on the device by ONE
launch per batch (preprocess.LinePreprocessor, kernel K7, instead of cv2 + albumentations + a copy per image) -- and,
in addition, preprocessed tensors [3, img_h, img_w] in [-1, 1] (or an already batched [B, 3, H, W] tensor).

The backbone runs as ``model.FoldedBackbone`` (BatchNorm folded into the convolutions, fused conv + bias + ReLU, the SE
tails on this library's kernels): ``backbone_dtype=torch.float32`` (default) equals the module to rounding,
``torch.bfloat16`` is the fast setting, ``None`` keeps the module's own eager path."""
from __future__ import annotations

from typing import List, Union

import torch

from .charset import ctc_alphabet, decode_tokens, load_charset
from .decode import _to_host, ctc_greedy_ids, ids_to_text_host
from .model import RCNN
from .preprocess import LinePreprocessor


class OCRInference:
    def __init__(self, model_path=None, charset_path=None, device: str = "auto", img_h: int = 64,
                 img_w: int = 256, model: RCNN | None = None, hidden_size: int = 256,
                 decoder: str | None = None, backbone_dtype: torch.dtype | None = torch.float32):
        if device == "auto":
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("rcnn-ocr_b200 has no CPU path; pass a CUDA device")
        self.img_h, self.img_w = img_h, img_w
        self.itos, self.stoi = load_charset(charset_path)
        self.alphabet, self.num_ctc_classes, self.blank = ctc_alphabet(self.itos)
        self.pad_id, self.eos_id = self.stoi.get("<PAD>", 0), self.stoi.get("<EOS>", 2)
        self.blank_id = self.stoi.get("<BLANK>")
        self.unused_checkpoint_keys: List[str] = []
        if model is None:
            state = None
            if model_path is not None:
                state = torch.load(model_path, map_location="cpu")
                if isinstance(state, dict) and "model_state" in state:
                    state = state["model_state"]
                elif isinstance(state, dict) and "model_state_dict" in state:
                    state = state["model_state_dict"]
                has_ctc = any(k.startswith("ctc_head.") for k in state)
                has_attn = any(k.startswith("attn.") for k in state)
                if decoder is None:
                    # a reference checkpoint carries attn.* and no CTC head: decode it the reference's way
                    decoder = "ctc" if has_ctc or not has_attn else "attention"
                elif decoder == "ctc" and not has_ctc:
                    raise ValueError(f"{model_path} has no ctc_head.* weights: its predictions through a randomly "
                                     "initialised CTC head would be garbage; pass decoder='attention' (or None) "
                                     "to decode a reference checkpoint with its own attention decoder")
            decoder = decoder or "ctc"
            model = RCNN(num_classes=len(self.itos), hidden_size=hidden_size,
                         sos_id=self.stoi.get("<SOS>", 1), eos_id=self.eos_id,
                         pad_id=self.pad_id, blank_id=self.blank_id, decoder=decoder)
            if state is not None:
                if decoder == "ctc":
                    model.load_state_dict(state, strict=True)
                else:
                    self.unused_checkpoint_keys = model.load_reference_state_dict(state)
                    if self.unused_checkpoint_keys:
                        import warnings
                        warnings.warn(f"{len(self.unused_checkpoint_keys)} checkpoint keys were not used: "
                                      f"{self.unused_checkpoint_keys[:5]}")
        self.model = model.to(self.device).eval()
        # inference-only copy of the backbone with BatchNorm folded into the convolutions and conv + bias + ReLU as one
        # cuDNN call (model.FoldedBackbone).  float32 (default) reproduces the module to rounding; torch.bfloat16 is the fast
        # setting bench.py's cfg 1 measures; None keeps the module as it is
        self.model.fold_backbone(backbone_dtype)
        self.transform = LinePreprocessor(img_h, img_w, self.device)      # get_val_transform(img_h, img_w), on the device

    @torch.no_grad()
    def predict(self, images, max_length: int = 25, batch_size: int = 32, return_confidence: bool = False):
        """inference.py:126-194: one item (path / PIL image / numpy array / [3,H,W] tensor) -> ``str`` (or
        ``(str, confidence)``); a list (or a batched [B,3,H,W] tensor) -> a list."""
        if isinstance(images, torch.Tensor):
            is_single = images.dim() == 3
            items = [images] if is_single else list(images)
        else:
            is_single = not isinstance(images, list)
            items = [images] if is_single else list(images)
        results = []
        for i in range(0, len(items), batch_size):
            part = items[i:i + batch_size]
            if all(isinstance(t, torch.Tensor) for t in part):
                chunk = torch.stack([t.float() for t in part])
                if not chunk.is_cuda:
                    chunk = chunk.pin_memory().to(self.device, non_blocking=True)
            elif any(isinstance(t, torch.Tensor) for t in part):
                raise ValueError("a batch mixes preprocessed tensors with raw images")
            else:
                chunk = self.transform(part)          # decode on the host, everything else in one launch
            logits = self.model(chunk, is_train=False, batch_max_length=max_length)
            if getattr(self.model, "attn", None) is not None:     # the reference's decode step, inference.py:167-189
                pred = logits.argmax(dim=-1)
                if return_confidence:
                    maxp = torch.softmax(logits, dim=-1).max(dim=-1)[0]
                    valid = (pred != self.pad_id) & (pred != self.eos_id)
                    conf = torch.where(valid.sum(1) > 0, (maxp * valid).sum(1) / valid.sum(1).clamp(min=1),
                                       torch.zeros_like(maxp[:, 0])).cpu().tolist()
                for j, row in enumerate(pred.cpu()):
                    text = decode_tokens(row, self.itos, pad_id=self.pad_id, eos_id=self.eos_id, blank_id=self.blank_id)
                    results.append((text, conf[j]) if return_confidence else text)
                continue
            out = ctc_greedy_ids(logits, blank=self.blank, return_confidence=return_confidence)
            ids_h, lens_h = _to_host(out[0], out[1])
            conf_h = out[2].cpu().tolist() if return_confidence else None
            texts, _ = ids_to_text_host(ids_h, lens_h, self.alphabet)
            for j, text in enumerate(texts):
                results.append((text, conf_h[j]) if return_confidence else text)
        return results[0] if is_single else results
