"""Pin the CPU oracle (oracle/) against the golden vectors generated from the
reference's own code and torch float64 by tests/make_golden.py.  CPU only."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import host_ref, lstm_ref
from conftest import GOLDEN, golden


def _names(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


@pytest.mark.parametrize("name", _names("decodefn_"))
def test_decode_function_goldens(name):
    """The reference's decode() = log_softmax + greedy; the oracle's greedy on the raw scores gives the same
    strings (log_softmax is argmax-invariant on these inputs)."""
    d = golden(name)
    alphabet = json.loads(str(d["alphabet"]))
    texts, seqs = oracle.ctc_greedy_decoder(d["logits"], alphabet, blank=0)
    assert seqs == json.loads(str(d["seqs"])) and texts == json.loads(str(d["texts"]))


@pytest.mark.parametrize("name", _names("decode_"))
def test_decode_matches_reference(name):
    d = golden(name)
    alphabet = json.loads(str(d["alphabet"]))
    blank = int(d["blank"]) if "blank" in d.files else 0
    texts, seqs = oracle.ctc_greedy_decoder(d["logits"], alphabet, blank=blank)
    assert seqs == json.loads(str(d["seqs"]))
    assert texts == json.loads(str(d["texts"]))


@pytest.mark.parametrize("name", _names("ctc_"))
def test_ctc_matches_torch_f64(name):
    d = golden(name)
    x, tg, il, tl = d["x"], d["targets"], d["input_lengths"], d["target_lengths"]
    blank = int(d["blank"])
    for zi in (0, 1):
        nll, _ = oracle.ctc_nll_and_grad(x, tg, il, tl, blank, bool(zi), from_logits=True,
                                         want_grad=False)
        np.testing.assert_allclose(nll, d[f"nll_zi{zi}"], rtol=1e-12, atol=1e-12)
        for red in ("mean", "sum"):
            if f"grad_{red}_zi{zi}" not in d.files:
                continue
            loss, grad = oracle.ctc_loss(x, tg, il, tl, blank, red, bool(zi), from_logits=True)
            np.testing.assert_allclose(loss, d[f"loss_{red}_zi{zi}"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(grad, d[f"grad_{red}_zi{zi}"], rtol=1e-9, atol=1e-11,
                                       equal_nan=True)
    # 1-D concatenated targets give the same result as padded targets
    loss_c, grad_c = oracle.ctc_loss(x, d["targets_concat"], il, tl, blank, "mean", True)
    np.testing.assert_allclose(loss_c, d["loss_mean_zi1"], rtol=1e-12)
    np.testing.assert_allclose(grad_c, d["grad_mean_zi1"], rtol=1e-9, atol=1e-11)


def test_ctc_logprob_convention_matches_aten():
    """from_logits=False returns exp(lp) - occupancy, what ATen's backward emits."""
    d = golden("ctc_small.npz")
    x = torch.tensor(d["x"])
    lp = x.log_softmax(2).detach().requires_grad_(True)
    tg, il, tl = (torch.tensor(d[k]) for k in ("targets", "input_lengths", "target_lengths"))
    loss = torch.nn.functional.ctc_loss(lp, tg, il, tl, reduction="sum", zero_infinity=True)
    g, = torch.autograd.grad(loss, lp)
    _, grad = oracle.ctc_loss(lp.detach().numpy(), d["targets"], d["input_lengths"],
                              d["target_lengths"], 0, "sum", True, from_logits=False)
    np.testing.assert_allclose(grad, g.numpy(), rtol=1e-9, atol=1e-11)


def test_ctc_bruteforce_tiny():
    """Enumerate every alignment for T<=5, C<=3 and compare the likelihood."""
    import itertools
    rng = np.random.default_rng(0)
    for T, C, tgt in [(3, 3, [1]), (4, 3, [1, 2]), (5, 3, [1, 1]), (4, 2, [1, 1]), (2, 3, [])]:
        x = rng.normal(size=(T, 1, C))
        lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
        total = 0.0
        for path in itertools.product(range(C), repeat=T):
            col, prev = [], 0
            for p in path:
                if p != 0 and p != prev:
                    col.append(p)
                prev = p
            # collapse with blank separation: repeated labels need a blank in between
            out, prev = [], None
            for p in path:
                if p != prev and p != 0:
                    out.append(p)
                prev = p
            if out == tgt:
                total += np.exp(sum(lp[t, 0, p] for t, p in enumerate(path)))
        nll, _ = oracle.ctc_nll_and_grad(x, np.array([tgt + [0] * (2 - len(tgt))]), [T], [len(tgt)],
                                         0, False, True, want_grad=False)
        expect = -np.log(total) if total > 0 else np.inf
        np.testing.assert_allclose(nll[0], expect, rtol=1e-10)


@pytest.mark.parametrize("name", _names("bilstm_"))
def test_bilstm_forward_matches_reference(name):
    d = golden(name)
    params = {k[2:]: d[k] for k in d.files if k.startswith("p.")}
    out = oracle.bilstm_forward(d["x"], params)
    np.testing.assert_allclose(out, d["out"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", _names("bilstm_"))
def test_lstm_ref_forward_backward_matches_reference(name):
    d = golden(name)
    params = {k[2:]: torch.tensor(d[k], requires_grad=True) for k in d.files if k.startswith("p.")}
    x = torch.tensor(d["x"], requires_grad=True)
    out = lstm_ref.bilstm_block(x, params)
    np.testing.assert_allclose(out.detach().numpy(), d["out"], rtol=1e-10, atol=1e-12)
    (out * torch.tensor(d["w"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), d["gx"], rtol=1e-8, atol=1e-11)
    for k, v in params.items():
        np.testing.assert_allclose(v.grad.numpy(), d["g." + k], rtol=1e-8, atol=1e-10, err_msg=k)


def test_enc_rnn_stack_matches_reference():
    d = golden("encrnn_small.npz")
    params = {k[2:]: torch.tensor(d[k], requires_grad=True) for k in d.files if k.startswith("p.")}
    x = torch.tensor(d["x"], requires_grad=True)
    out = lstm_ref.enc_rnn(x, params)
    np.testing.assert_allclose(out.detach().numpy(), d["out"], rtol=1e-10, atol=1e-12)
    (out * torch.tensor(d["w"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), d["gx"], rtol=1e-8, atol=1e-11)
    for k, v in params.items():
        np.testing.assert_allclose(v.grad.numpy(), d["g." + k], rtol=1e-8, atol=1e-10, err_msg=k)


def test_host_ref_charset_and_metrics(tmp_path):
    p = tmp_path / "cs.txt"
    p.write_text("<PAD>\n<SOS>\n<EOS>\n \na\n\nb\n", encoding="utf-8")
    itos, stoi = host_ref.load_charset(str(p))
    assert itos == ["<PAD>", "<SOS>", "<EOS>", " ", "a", "b"] and stoi["a"] == 4
    assert host_ref.decode_tokens([4, 0, 5, 2, 4], itos, 0, 2) == "ab"
    assert host_ref.character_error_rate("kitten", "sitting") == pytest.approx(3 / 6)
    assert host_ref.character_error_rate("", "x") == float("inf")
    assert host_ref.character_error_rate("", "") == 0.0
    assert host_ref.compute_accuracy(["a", "b"], ["a", "c"]) == 0.5
    assert host_ref.compute_accuracy([], []) == 0.0
    # jiwer's documented example: one substitution in four reference words
    assert host_ref.word_error_rate("hello world how are", "hello duck how are") == 0.25
    assert host_ref.word_error_rate("  a   b ", "a b") == 0.0           # whitespace runs collapse, ends are stripped
    assert host_ref.word_error_rate("a b", "a b c d") == 1.0            # insertions can push WER past 1


def test_attention_port_matches_reference_module_outputs():
    """oracle/ref_port.RefAttention (the CPU baseline of the attention decoder) against the reference module's
    own greedy outputs (tests/golden/attn_*.npz, made by importing model/model.py in the build container)."""
    import glob, os
    import numpy as np
    import torch
    from conftest import GOLDEN
    from oracle.ref_port import RefAttention
    n = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "attn_*.npz"))):
        if os.path.basename(path).startswith("attn_train_"):
            continue                       # gradients of the training path: tests/test_attention_gpu.py
        d = np.load(path)
        if "seed_scale" in d.files:
            continue
        B, T, C, H, V, steps, blank = [int(v) for v in d["dims"]]
        m = RefAttention(C, H, V, sos_id=1, blank_id=None if blank < 0 else blank).eval()
        m.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}, strict=True)
        got = m.greedy(torch.from_numpy(d["batch_H"]), steps - 1).numpy()
        np.testing.assert_allclose(got, d["probs"], rtol=1e-5, atol=1e-5)
        n += 1
    assert n >= 2


def test_cnn_adapter_matches_the_reference_backbone():
    """rcnn_ocr_b200.SEResNet31 (the state-dict compatible restatement that RCNN keeps so that it stays a drop-in,
    and that oracle/ref_port.RefRCNN times on the CPU) against the reference's own module (model/seresnet31.py:70-187,
    imported by tests/make_golden.py): same keys, same eval-mode output, same training-mode output and gradients."""
    import rcnn_ocr_b200 as R
    from conftest import recipe_state_dict
    d = golden("cnn_eval_2x32x64.npz")
    ours = R.SEResNet31(3, 512).eval()
    assert sorted(ours.state_dict().keys()) == json.loads(str(d["keys"]))
    ours.load_state_dict(recipe_state_dict(ours, int(d["seed"])), strict=True)
    with torch.no_grad():
        y = ours(torch.from_numpy(d["x"]))
    np.testing.assert_allclose(y.numpy(), d["y"], rtol=1e-5, atol=1e-5 * np.abs(d["y"]).max())
    t = golden("cnn_train_4x32x64.npz")
    ours.train()
    ours.load_state_dict(recipe_state_dict(ours, int(t["seed"])), strict=True)
    x = torch.from_numpy(t["x"]).requires_grad_(True)
    yt = ours(x)
    yt.square().mean().backward()
    np.testing.assert_allclose(yt.detach().numpy(), t["y"], rtol=1e-4, atol=1e-5 * np.abs(t["y"]).max())
    np.testing.assert_allclose(x.grad.numpy(), t["dx"], rtol=1e-3, atol=1e-4 * np.abs(t["dx"]).max())
    np.testing.assert_allclose(ours.conv0[0].weight.grad.numpy(), t["dw0"], rtol=1e-3, atol=1e-4 * np.abs(t["dw0"]).max())


@pytest.mark.parametrize("name", ["preproc_32x128", "preproc_64x256"])
def test_input_step_oracle_matches_the_reference_transform(name):
    """oracle/host_ref.resize_and_pad against canvases produced by the reference's own ResizeAndPadA.apply."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    ih, iw = int(d["img_h"]), int(d["img_w"])
    for i in range(int(d["n"])):
        got = host_ref.resize_and_pad(d[f"img{i}"], ih, iw)
        np.testing.assert_array_equal(got, d["canvas"][i], err_msg=f"image {i} {d[f'img{i}'].shape}")
    chw = host_ref.normalize_chw(d["canvas"][0])
    assert chw.shape == (3, ih, iw) and chw.dtype == np.float32 and chw.max() <= 1.0 and chw.min() >= -1.0


@pytest.mark.parametrize("name", ["pack_attn_ml25", "pack_attn_ml5"])
def test_target_packing_matches_the_reference(name):
    """The vectorised packers (product host code, no kernel involved) and the oracle restatement against the reference's
    pack_attention_targets; the CTC packer is the same character walk shifted by one class (class k = itos[k-1])."""
    from rcnn_ocr_b200 import preprocess
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    texts, ml = [str(t) for t in d["texts"]], int(d["max_len"])
    stoi = {str(t): i for i, t in enumerate(d["itos"])}      # configs/charset.txt as the generator read it
    for impl in (host_ref.pack_attention_targets, preprocess.pack_attention_targets):
        ti, ty, ln = [np.asarray(a) for a in impl(texts, stoi, ml)]
        np.testing.assert_array_equal(ti, d["text_in"])
        np.testing.assert_array_equal(ty, d["target_y"])
        np.testing.assert_array_equal(ln, d["lengths"])
    tg, tl = preprocess.pack_ctc_targets(texts, stoi, max_len=ml)
    np.testing.assert_array_equal(tl.numpy(), d["lengths"] - 1)
    want = np.concatenate([d["target_y"][i, :d["lengths"][i] - 1] + 1 for i in range(len(texts))])
    np.testing.assert_array_equal(tg.numpy(), want)
