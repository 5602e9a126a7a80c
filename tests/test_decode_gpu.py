"""GPU parity: greedy CTC decode kernel (K4) vs the oracle and the reference's golden outputs.
Bit-exact (integer/index work)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import oracle
import rcnn_ocr_b200 as R
from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu


def _names(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


@pytest.mark.parametrize("name", _names("decode_"))
def test_golden_reference_outputs(name):
    d = golden(name)
    alphabet = json.loads(str(d["alphabet"]))
    blank = int(d["blank"]) if "blank" in d.files else 0
    logits = torch.from_numpy(d["logits"]).cuda()
    texts, seqs = R.ctc_greedy_decoder(logits, alphabet, blank=blank)
    assert seqs == json.loads(str(d["seqs"]))
    assert texts == json.loads(str(d["texts"]))
    # host tensors take the same path through a pinned H2D copy
    texts_h, seqs_h = R.ctc_greedy_decoder(torch.from_numpy(d["logits"]), alphabet, blank=blank)
    assert seqs_h == seqs and texts_h == texts


@pytest.mark.parametrize("name", _names("decodefn_"))
def test_decode_function_matches_the_reference(name):
    """``decode(ctc_out, alphabet, method)`` against the reference's own ``decode`` (training/utils.py:153-162:
    tuple unwrap, log_softmax, greedy with the shape heuristic), bit-exact strings and index lists."""
    d = golden(name)
    alphabet = json.loads(str(d["alphabet"]))
    logits = torch.from_numpy(d["logits"]).cuda()
    out = (logits, torch.zeros(1, device="cuda")) if int(d["is_tuple"]) else logits
    texts, seqs = R.decode(out, alphabet, method="greedy")
    assert seqs == json.loads(str(d["seqs"]))
    assert texts == json.loads(str(d["texts"]))
    texts2, seqs2 = R.decode(out, alphabet)          # default method
    assert (texts2, seqs2) == (texts, seqs)
    with pytest.raises(ValueError):
        R.decode(out, alphabet, method="beam")


def _check(logits, blank=0, **kw):
    ids, lens = R.ctc_greedy_ids(logits, blank=blank, **kw)
    x = logits.float().cpu().numpy()
    if kw.get("batch_first", True) is False:
        x = x.transpose(1, 0, 2)
    eids, elens = oracle.ctc_greedy_ids(x, blank)
    np.testing.assert_array_equal(lens.cpu().numpy(), elens)
    np.testing.assert_array_equal(ids.cpu().numpy(), eids)


@pytest.mark.parametrize("B,T,C", [(1, 1, 1), (1, 1, 2), (3, 7, 5), (32, 16, 195), (256, 64, 195),
                                   (257, 33, 196), (5, 64, 1024), (2048, 16, 195), (20000, 8, 37),
                                   (7, 130, 195), (64, 64, 8)])
def test_random_vs_oracle(B, T, C):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    _check(torch.randn(B, T, C, device="cuda", generator=g))


def test_quantised_ties_and_blank_runs():
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(0, 3, (300, 40, 11), device="cuda", generator=g).float()
    _check(x)
    _check(x, blank=4)
    _check(torch.zeros(9, 12, 195, device="cuda"))           # all ties -> class 0 -> all blank
    y = torch.full((4, 10, 6), -1.0, device="cuda")
    y[:, :, 3] = 2.0                                          # one long run of a single label
    _check(y)


def test_nan_inf_rows():
    x = torch.randn(64, 20, 50, device="cuda")
    x[0, 0, 7] = float("nan"); x[1, 3, 0] = float("nan"); x[1, 3, 9] = float("nan")
    x[2, 1, 5] = float("inf"); x[2, 1, 8] = float("inf"); x[3] = float("-inf")
    x[4, 2, 30] = float("inf"); x[4, 2, 31] = float("nan"); x[5, :, :] = -0.0; x[5, :, 4] = 0.0
    _check(x)


def test_layouts_and_dtypes():
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(40, 24, 195, device="cuda", generator=g)
    _check(x.permute(1, 0, 2).contiguous(), batch_first=False)       # time-major storage
    _check(x.permute(1, 0, 2), batch_first=True)                      # strided batch-first view
    _check(x[:, :, 3:190])                                            # misaligned rows, C=187
    _check(x[:, 1::2])                                                # strided frames
    _check(x.bfloat16())
    _check(x.bfloat16()[:, :, 1:])
    _check(x.half())                                                  # converted to float32


def test_empty_inputs():
    ids, lens = R.ctc_greedy_ids(torch.zeros(0, 5, 7, device="cuda"))
    assert ids.shape == (0, 5) and lens.shape == (0,)
    ids, lens = R.ctc_greedy_ids(torch.zeros(3, 0, 7, device="cuda"))
    assert ids.shape == (3, 0) and lens.cpu().tolist() == [0, 0, 0]
    texts, seqs = R.ctc_greedy_decoder(torch.zeros(3, 0, 7, device="cuda"), "abcdef", batch_first=True)
    assert texts == ["", "", ""] and seqs == [[], [], []]


def test_confidence_output():
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(50, 32, 195, device="cuda", generator=g) * 3
    ids, lens, conf = R.ctc_greedy_ids(x, return_confidence=True)
    p = x.softmax(-1)
    mx, am = p.max(-1)
    valid = am != 0
    expect = torch.where(valid.sum(1) > 0, (mx * valid).sum(1) / valid.sum(1).clamp(min=1), torch.zeros(50, device="cuda"))
    torch.testing.assert_close(conf, expect, rtol=1e-4, atol=1e-6)
    eids, elens = oracle.ctc_greedy_ids(x.cpu().numpy(), 0)
    np.testing.assert_array_equal(ids.cpu().numpy(), eids)


def test_full_size_sweep_properties():
    """BASELINE config 5 sizes: decode is idempotent under re-encoding the argmax as one-hot
    logits, and every emitted id is a non-blank class with no adjacent... (oracle still checks)."""
    g = torch.Generator(device="cuda").manual_seed(77)
    x = torch.randn(4096, 64, 195, device="cuda", generator=g)
    ids, lens = R.ctc_greedy_ids(x)
    eids, elens = oracle.ctc_greedy_ids(x.cpu().numpy(), 0)
    np.testing.assert_array_equal(ids.cpu().numpy(), eids)
    onehot = torch.nn.functional.one_hot(x.argmax(-1), 195).float()
    ids2, lens2 = R.ctc_greedy_ids(onehot)
    assert torch.equal(ids, ids2) and torch.equal(lens, lens2)
    assert int(ids.max()) < 195 and not bool(((ids == 0)).any())
    pad = torch.arange(64, device="cuda")[None, :] >= lens[:, None]
    assert bool((ids[pad] == -1).all()) and bool((ids[~pad] > 0).all())


def test_deferred_host_half_equals_the_immediate_one():
    """ids_to_text_async: three batches in flight (own pinned buffers and events), results asked for in order and out of
    order, equal to ids_to_text of the same ids."""
    import rcnn_ocr_b200 as R
    g = torch.Generator(device="cuda").manual_seed(3)
    alphabet = [chr(0x430 + i) for i in range(40)]
    batches = [torch.randn(17, 23, 41, device="cuda", generator=g) for _ in range(3)]
    outs = [R.ctc_greedy_ids(b) for b in batches]
    outs = [(i.clone(), l.clone()) for i, l in outs]
    want = [R.ids_to_text(i, l, alphabet) for i, l in outs]
    pend = [R.ids_to_text_async(i, l, alphabet) for i, l in outs]
    for k in (1, 0, 2):
        assert pend[k].result() == want[k]
