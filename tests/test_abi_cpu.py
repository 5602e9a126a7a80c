"""CPU-only checks of the boundary: the shared library loads, exports every symbol that
include/rcnn_ocr_b200.h declares, validates arguments before touching CUDA, and the host-side
mirror of the reference interface behaves like the reference (errors included)."""
import os
import re

import pytest
import torch

import rcnn_ocr_b200 as R
from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rcnn_ocr_b200.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = R.lib()
    names = _declared_symbols()
    assert len(names) >= 6
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rcnn_ocr_b200.h but not exported"
    assert L.rcnn_version() >= 100


def test_argument_errors_do_not_need_a_gpu():
    L = R.lib()
    assert L.rcnn_ctc_greedy(None, 0, 4, 4, 0, 0, 0, 0, None, None, None, None) == 1
    assert b"bad shape" in L.rcnn_last_error()
    assert L.rcnn_ctc_greedy(None, 7, 4, 4, 5, 20, 5, 0, None, None, None, None) == 1
    assert b"dtype" in L.rcnn_last_error()
    # blank outside the class range
    assert L.rcnn_ctc_loss(None, 1, 4, 2, 5, 10, 5, None, 0, None, None, 3, 9, 1, 0,
                           None, None, None, 0, 0, None, 0, None) == 1
    # target length beyond the supported lattice
    assert L.rcnn_ctc_loss(None, 1, 4, 2, 5, 10, 5, None, 0, None, None, 300, 0, 1, 0,
                           None, None, None, 0, 0, None, 0, None) == 1


def test_new_entry_points_validate_arguments():
    L = R.lib()
    assert L.rcnn_edit_distance(None, 0, None, None, None, None, 4, None, None, 5, 0, None, None, None, None) == 1
    assert b"null pointer" in L.rcnn_last_error()
    assert L.rcnn_edit_distance(None, 0, None, None, None, None, 0, None, None, 5, 0, None, None, None, None) == 0
    assert L.rcnn_lstm_backward_workspace_bytes(256, 64, 512) == 2 * 8 * 16 * 16 * 4096
    assert L.rcnn_lstm_backward_workspace_bytes(256, 64, 100) == 0


def test_workspace_sizes():
    L = R.lib()
    small = L.rcnn_ctc_workspace_bytes(64, 256, 195, 32)
    assert small >= 8 * 257 and small < 16384          # offsets only: tables fit shared memory
    big = L.rcnn_ctc_workspace_bytes(512, 4, 195, 200)  # lattice spills to the workspace
    assert big >= 3 * 512 * 401 * 4 * 4


def test_no_cpu_fallback():
    x = torch.randn(4, 2, 5).log_softmax(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        R.ctc_loss(x, torch.tensor([[1], [2]]), [4, 4], [1, 1])
    with pytest.raises(ValueError):
        R.ctc_loss(x, torch.tensor([[1], [2]]), [4, 4], [1, 1], reduction="avg")
    with pytest.raises(ValueError, match="Unsupported decode method"):
        R.decode(torch.zeros(2, 2, 3), "ab", method="beam")


def test_charset_matches_reference_semantics(tmp_path):
    p = tmp_path / "charset.txt"
    p.write_text("<PAD>\n<SOS>\n<EOS>\n \na\nb\n\nc", encoding="utf-8")
    itos, stoi = R.load_charset(str(p))
    assert itos == ["<PAD>", "<SOS>", "<EOS>", " ", "a", "b", "c"]
    assert stoi["<PAD>"] == 0 and stoi[" "] == 3 and stoi["c"] == 6
    alphabet, C, blank = R.ctc_alphabet(itos)
    assert C == 8 and blank == 0 and alphabet[stoi["a"]] == "a"
    assert R.decode_tokens([4, 0, 5, 2, 6], itos, 0, 2) == "ab"
    from rcnn_ocr_b200.charset import encode_ctc_targets
    flat, lens = encode_ctc_targets(["ab", "", "c a"], stoi)
    assert flat == [5, 6, 7, 4, 5] and lens == [2, 0, 3]


def test_host_half_of_the_decode_matches_the_per_line_mapping():
    """ids_to_text_host (batch-wide gather + one utf-32 decode for single-character alphabets, object table
    otherwise) against the reference's per-line ``''.join(alphabet[p - 1] ...)`` (training/utils.py:146): ragged
    and empty lines, an empty batch, multi-character tokens, non-BMP characters, out-of-range ids."""
    import numpy as np
    from rcnn_ocr_b200.decode import ids_to_text_host
    rng = np.random.default_rng(0)
    for alphabet in ([chr(0x430 + i) for i in range(60)], [chr(0x4E00 + i) for i in range(194)],
                     ["a", "<PAD>", "bc", "\U0001F600", " "], ["x"]):
        C = len(alphabet)
        for B, T in ((0, 7), (1, 1), (5, 0), (37, 64), (256, 64)):
            lens = rng.integers(0, T + 1, B).astype(np.int32)
            ids = np.full((B, T), -1, np.int32)
            for b in range(B):
                ids[b, :lens[b]] = rng.integers(1, C + 1, lens[b])
            texts, seqs = ids_to_text_host(ids, lens, alphabet)
            want_seqs = [ids[b, :lens[b]].tolist() for b in range(B)]
            want_texts = ["".join(alphabet[k - 1] for k in row) for row in want_seqs]
            assert seqs == want_seqs and texts == want_texts
    with pytest.raises(IndexError):
        ids_to_text_host(np.array([[3, -1]], np.int32), np.array([1], np.int32), ["a", "b"])


def test_encoder_container_keeps_the_reference_state_dict_keys():
    """make_enc_rnn returns an nn.Sequential subclass (EncRNN: weight-conversion prefetch in training) whose
    state-dict keys and shapes are those of the reference's enc_rnn (model/model.py:195-198); the hand-off
    format to the CTC head is a constructor choice, not a parameter."""
    import torch.nn as nn
    enc = R.make_enc_rnn(512, 64)
    assert isinstance(enc, nn.Sequential) and len(enc) == 2
    want = []
    for i, isz in ((0, 512), (1, 64)):
        for sfx in ("", "_reverse"):
            want += [(f"{i}.rnn.weight_ih_l0{sfx}", (256, isz)), (f"{i}.rnn.weight_hh_l0{sfx}", (256, 64)),
                     (f"{i}.rnn.bias_ih_l0{sfx}", (256,)), (f"{i}.rnn.bias_hh_l0{sfx}", (256,))]
        want += [(f"{i}.linear.weight", (64, 128)), (f"{i}.linear.bias", (64,))]
    got = [(k, tuple(v.shape)) for k, v in enc.state_dict().items()]
    assert got == want
    assert enc[0].out_dtype == torch.bfloat16 and enc[1].out_dtype == torch.float32
    assert R.make_enc_rnn(512, 64, out_dtype=torch.bfloat16)[1].out_dtype == torch.bfloat16
    from rcnn_ocr_b200 import ops
    assert ops.weight_grads_supported(512, 512, 64) and ops.weight_grads_supported(256, 256, 128)
    assert not ops.weight_grads_supported(512, 512, 16) and not ops.weight_grads_supported(64, 64, 64)
    assert ops.fused_forward_supported(512, 512) and not ops.fused_forward_supported(576, 512)
