"""rcnn-ocr_b200: B200-native drop-in for RCNN-OCR's sequence-recognition hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the arithmetic
runs in hand-written sm_100a CUDA kernels behind the C ABI in include/rcnn_ocr_b200.h
(loaded with ctypes from librcnn_ocr_b200.so).  There is no CPU fallback: every op raises
if the library is missing or the tensors are not on a CUDA device.

Surface (mirrors the reference; citations are path:line under the reference tree):
  BidirectionalLSTM, RCNN        model/model.py:151-163, 166-227
  Attention (inference)          model/model.py:23-148
  CTCLoss, ctc_loss              nn.CTCLoss call site at training/train.py:289,503-505
  ctc_greedy_decoder, decode     training/utils.py:122-162
  load_charset, decode_tokens    data/transforms.py:39-59, 196-206
  validation_metrics, ...        training/metrics.py:5-32 as used at training/train.py:582-598
  LinePreprocessor, pack_*       data/transforms.py:62-157,179 (ResizeAndPadA + Normalize + ToTensorV2, target packing)
"""
from ._lib import lib, library_path, LibraryMissing  # noqa: F401
from .charset import load_charset, decode_tokens, ctc_alphabet  # noqa: F401
from .decode import ctc_greedy_decoder, decode, ctc_greedy_ids, ids_to_text, ids_to_text_async, PendingTexts  # noqa: F401
from .ctc import CTCLoss, ctc_loss, ctc_loss_from_logits  # noqa: F401
from .model import BidirectionalLSTM, CTCHead, FoldedBackbone, RCNN, SEResNet31, make_enc_rnn  # noqa: F401
from .inference import OCRInference  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .attention import Attention, AttentionCell  # noqa: F401
from .preprocess import LinePreprocessor, pack_ctc_targets, pack_attention_targets  # noqa: F401
from .metrics import CharsetTable, edit_stats, character_error_rates, word_error_rates, validation_metrics  # noqa: F401

__version__ = "0.1.0"
