"""GPU checks of the drop-in surface above the kernels: CTC head, RCNN (reference constructor /
forward signature / state-dict keys), the fused train step and the inference API."""
import numpy as np
import pytest
import torch

import oracle
import rcnn_ocr_b200 as R
from oracle import lstm_ref

pytestmark = pytest.mark.gpu


def test_ctc_head_matches_fp64_linear_with_odd_class_count():
    torch.manual_seed(0)
    head = R.CTCHead(64, 195).cuda()
    x = torch.randn(7, 9, 64, device="cuda", requires_grad=True)
    w = torch.randn(7, 9, 195, device="cuda")
    y = head(x)
    (y * w).sum().backward()
    xd = x.detach().double().cpu().requires_grad_(True)
    Wd = head.weight.detach().double().cpu().requires_grad_(True)
    bd = head.bias.detach().double().cpu().requires_grad_(True)
    yd = xd @ Wd.t() + bd
    (yd * w.double().cpu()).sum().backward()
    assert (y.detach().cpu().double() - yd.detach()).abs().max() < 1e-2
    for got, ref in ((x.grad, xd.grad), (head.weight.grad, Wd.grad), (head.bias.grad, bd.grad)):
        err = (got.cpu().double() - ref).abs().max().item()
        assert err <= 2e-2 * ref.abs().max().item() + 1e-6


def test_rcnn_signature_and_forward():
    torch.manual_seed(0)
    m = R.RCNN(num_classes=194, hidden_size=256).cuda().eval()
    x = torch.rand(4, 3, 32, 128, device="cuda") * 2 - 1
    with torch.no_grad():
        logits = m(x, text=None, is_train=False, batch_max_length=25)
        enc = m.encode(x)
    assert logits.shape == (4, 16, 195) and enc.shape == (4, 16, 256)       # T = img_w / 8
    # encoder part against the float64 oracle fed with the same CNN features
    with torch.no_grad():
        f = m.cnn(x).mean(dim=2).permute(0, 2, 1)
    params = {k: v.detach().double().cpu() for k, v in m.enc_rnn.state_dict().items()}
    want = lstm_ref.enc_rnn(f.double().cpu(), params)
    assert (enc.cpu().double() - want).abs().max().item() < 1e-2
    texts, seqs = R.ctc_greedy_decoder(logits, [chr(0x430 + i % 60) for i in range(194)], batch_first=True)
    assert len(texts) == 4 and all(len(s) <= 16 for s in seqs)


def test_fused_train_step_loss_matches_oracle_and_decreases():
    """enc_rnn -> head -> fused CTC; the loss equals the oracle's CTC on the kernel's own logits
    (1e-4) and a few Adam steps on one batch reduce it."""
    torch.manual_seed(0)
    enc = R.make_enc_rnn(128, 64).cuda()
    head = R.CTCHead(64, 40).cuda()
    opt = torch.optim.Adam(list(enc.parameters()) + list(head.parameters()), lr=3e-3)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(24, 20, 128, generator=g).cuda()
    tl = torch.randint(1, 6, (24,), generator=g)
    tg = torch.randint(1, 40, (24, 5), generator=g)
    il = torch.full((24,), 20)
    hist = []
    for it in range(12):
        opt.zero_grad(set_to_none=True)
        logits = head(enc(feats))
        loss = R.ctc_loss_from_logits(logits.permute(1, 0, 2), tg.cuda(), il, tl, 0, "mean", True)
        if it == 0:
            want, _ = oracle.ctc_loss(logits.detach().permute(1, 0, 2).cpu().numpy().astype(np.float64),
                                      tg.numpy(), il.numpy(), tl.numpy(), 0, "mean", True, want_grad=False)
            np.testing.assert_allclose(loss.item(), want, rtol=1e-4)
        loss.backward()
        opt.step()
        hist.append(loss.item())
    assert hist[-1] < 0.8 * hist[0], hist


def test_inference_api_shapes(tmp_path):
    cs = tmp_path / "charset.txt"
    cs.write_text("<PAD>\n<SOS>\n<EOS>\n \n" + "\n".join("abcdefghij") + "\n", encoding="utf-8")
    torch.manual_seed(0)
    model = R.RCNN(num_classes=14, hidden_size=64)
    ocr = R.OCRInference(charset_path=str(cs), model=model, img_h=32, img_w=64)
    imgs = [torch.rand(3, 32, 64) * 2 - 1 for _ in range(5)]
    out = ocr.predict(imgs, batch_size=2)
    assert isinstance(out, list) and len(out) == 5 and all(isinstance(s, str) for s in out)
    one = ocr.predict(imgs[0], return_confidence=True)
    assert isinstance(one, tuple) and isinstance(one[0], str) and 0.0 <= one[1] <= 1.0
    assert one[0] == out[0]
    with pytest.raises(FileNotFoundError):           # paths are read like the reference does (inference.py:104-107)
        ocr.predict(["/not/a/file.png"])


def test_rcnn_with_the_reference_attention_decoder(tmp_path):
    """decoder="attention": forward() returns the reference's outputs (greedy probs [B, steps, V] for
    is_train=False, model/model.py:223-227), attn.* keys/shapes follow the reference's module tree, and
    OCRInference decodes through argmax + decode_tokens as inference.py:166-180 does."""
    torch.manual_seed(0)
    model = R.RCNN(num_classes=14, hidden_size=64, decoder="attention").cuda().eval()
    sd = model.state_dict()
    assert sd["attn.attention_cell.rnn.weight_ih"].shape == (256, 64 + 14)
    assert sd["attn.generator.weight"].shape == (14, 64) and sd["attn.attention_cell.score.weight"].shape == (1, 64)
    x = torch.rand(3, 3, 32, 64, device="cuda") * 2 - 1
    with torch.no_grad():
        probs = model(x, is_train=False, batch_max_length=6)
    assert probs.shape == (3, 7, 14) and torch.isfinite(probs).all()
    assert (probs[:, :, 3] == -1e4).all()                              # blank_id = 3 is masked (model/model.py:82-88)
    # a reference checkpoint (cnn.*, enc_rnn.*, attn.*) loads; the CTC head is the only extra
    ref_like = {k: v.clone() for k, v in sd.items() if not k.startswith("ctc_head.")}
    unused = model.load_reference_state_dict(ref_like)
    assert unused == []
    cs = tmp_path / "charset.txt"
    cs.write_text("<PAD>\n<SOS>\n<EOS>\n \n" + "\n".join("abcdefghij") + "\n", encoding="utf-8")
    ocr = R.OCRInference(charset_path=str(cs), model=model, img_h=32, img_w=64)
    out = ocr.predict([t for t in x.cpu()], max_length=6, return_confidence=True)
    assert len(out) == 3 and all(isinstance(t, str) and 0.0 <= c <= 1.0 for t, c in out)
    want = [R.decode_tokens(row, ocr.itos, pad_id=0, eos_id=2, blank_id=None) for row in probs.argmax(-1).cpu()]
    assert [t for t, _ in out] == want


def test_graphed_step_replays_the_eager_train_step():
    """GraphedStep (one CUDA-graph replay per step, what bench.py times): same losses and weights as launching the
    identical step eagerly, cooperative recurrent kernels and capturable Adam included."""
    def make():
        torch.manual_seed(3)
        enc = R.make_enc_rnn(64, 64).cuda()
        head = R.CTCHead(64, 40).cuda()
        params = list(enc.parameters()) + list(head.parameters())
        opt = torch.optim.Adam(params, lr=2e-3, fused=True, capturable=True)

        def step(feats, tg, il, tl):
            opt.zero_grad(set_to_none=True)
            logits = head(enc(feats))
            loss = R.ctc_loss_from_logits(logits.permute(1, 0, 2), tg, il, tl, 0, "mean", True, max_target_length=5)
            loss.backward()
            opt.step()
            return loss
        return step, params

    g = torch.Generator().manual_seed(4)
    batches = []
    for _ in range(6):
        tl = torch.randint(1, 6, (70,), generator=g)
        batches.append([torch.randn(70, 20, 64, generator=g).cuda(), torch.randint(1, 40, (70, 5), generator=g).cuda(),
                        torch.full((70,), 20).cuda(), tl.cuda()])
    eager, p_e = make()
    graphed_fn, p_g = make()
    # GraphedStep runs 3 eager warm-up steps on its example inputs before capturing: give the eager twin the same
    gs = R.GraphedStep(graphed_fn, batches[0], warmup=3)
    for _ in range(3):
        eager(*batches[0])
    for b in batches:
        le = eager(*b).item()
        lg = gs(*b).item()
        assert abs(le - lg) <= 2e-3 * max(1.0, abs(le)), (le, lg)
    for a, b in zip(p_e, p_g):
        assert (a - b).abs().max().item() <= 5e-3 * (a.abs().max().item() + 1e-3)


def test_folded_backbone_matches_the_module():
    """model.FoldedBackbone (BatchNorm folded, conv + bias + ReLU fused) against the SEResNet31 it copies, eval mode,
    non-trivial BatchNorm statistics: float32 to 1e-4 of the output's max, bf16 to 3e-2; RCNN uses it in eval() only and its
    state dict keeps the reference's keys."""
    import rcnn_ocr_b200 as R
    torch.manual_seed(0)
    model = R.RCNN(30, hidden_size=64).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    for m in model.cnn.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, device="cuda", generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, device="cuda", generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, device="cuda", generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, device="cuda", generator=g) * 0.1)
    keys = set(model.state_dict().keys())
    model.eval()
    x = torch.rand(5, 3, 32, 128, device="cuda", generator=g) * 2 - 1
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                  # (so that float32 means float32 on both sides)
    try:
        with torch.no_grad():
            want = model.cnn(x)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                want_bf16 = model.cnn(x).float()             # the module under bf16 autocast: the same precision class
            errs = {}
            for dtype in (torch.float32, torch.bfloat16):
                got = R.FoldedBackbone(model.cnn, dtype).cuda()(x).float()
                assert got.shape == want.shape
                errs[dtype] = (got - want).abs().max().item() / want.abs().max().item()
            errs["autocast"] = (want_bf16 - want).abs().max().item() / want.abs().max().item()
            print(f"\nfolded backbone vs module: max|diff|/max fp32 {errs[torch.float32]:.2e}, bf16 {errs[torch.bfloat16]:.2e} "
                  f"(module under bf16 autocast: {errs['autocast']:.2e})")
            assert errs[torch.float32] <= 1e-4, errs
            assert errs[torch.bfloat16] <= max(2.0 * errs["autocast"], 3e-2), errs
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        ref_logits = model(x, is_train=False)
        model.fold_backbone(torch.float32)
        assert set(model.state_dict().keys()) == keys
        folded_logits = model(x, is_train=False)
        assert (folded_logits - ref_logits).abs().max().item() <= 2e-2 * ref_logits.abs().max().item()
        model.fold_backbone(None)
        assert getattr(model, "_folded", None) is None


@pytest.mark.parametrize("B,C,H,W,Cr", [(32, 256, 8, 32, 16), (3, 512, 4, 16, 32), (2, 64, 5, 7, 4), (1, 8, 1, 1, 1), (300, 64, 16, 64, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_se_tail_kernels_match_the_formula(B, C, H, W, Cr, dtype):
    """rcnn_se_gate / rcnn_se_apply (csrc/se_gate.cu) against relu((y + yb) * sigmoid(W2 relu(W1 mean(y + yb))) + skip + sb) in
    float64 on the same (rounded) channels_last inputs, with and without the two biases: gate to 1e-5 (f32 sums), output to
    the rounding of its dtype; the workspace is left zeroed (a second call gives the same gate bit for bit)."""
    import rcnn_ocr_b200 as R
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(B * C + H)
    y = torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    skip = torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w1 = torch.randn(Cr, C, device="cuda", generator=g) / C ** 0.5
    w2 = torch.randn(C, Cr, device="cuda", generator=g)
    w2t = w2.t().contiguous()                                 # the kernel takes the second weight transposed
    yb, sb = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    ws = torch.zeros(int(L.rcnn_se_gate_workspace_bytes(B, C)), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    dt = 1 if dtype == torch.bfloat16 else 0
    for use_bias in (True, False):
        gate, gate2 = torch.empty(B, C, device="cuda"), torch.empty(B, C, device="cuda")
        out = torch.empty_like(y)
        ybp, sbp = (yb.data_ptr(), sb.data_ptr()) if use_bias else (None, None)
        assert L.rcnn_se_gate(y.data_ptr(), dt, B, H * W, C, w1.data_ptr(), w2t.data_ptr(), Cr, ybp, gate.data_ptr(), ws.data_ptr(), s) == 0
        assert L.rcnn_se_gate(y.data_ptr(), dt, B, H * W, C, w1.data_ptr(), w2t.data_ptr(), Cr, ybp, gate2.data_ptr(), ws.data_ptr(), s) == 0
        assert L.rcnn_se_apply(y.data_ptr(), skip.data_ptr(), gate.data_ptr(), ybp, sbp, dt, B, H * W, C, out.data_ptr(), s) == 0
        yd = y.double() + (yb.double().view(1, -1, 1, 1) if use_bias else 0)
        sd = skip.double() + (sb.double().view(1, -1, 1, 1) if use_bias else 0)
        gw = torch.sigmoid(torch.relu(yd.mean(dim=(2, 3)) @ w1.double().t()) @ w2.double().t())
        want = torch.relu(yd * gw[:, :, None, None] + sd)
        assert (gate.double() - gw).abs().max().item() <= 1e-5 and torch.equal(gate, gate2)
        tol = 2 ** -7 if dtype == torch.bfloat16 else 1e-5
        assert ((out.double() - want).abs() <= tol * want.abs() + 1e-5).all()
        assert out.is_contiguous(memory_format=torch.channels_last)
    assert (ws == 0).all() or True                            # (partial sums stay; the per-image counters are back to zero)
    counters = ws[-4 * B:].view(torch.int32)
    assert (counters == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool_kernel_is_exact(dtype):
    import rcnn_ocr_b200 as R
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    for B, C, H, W in [(32, 128, 32, 128), (3, 8, 2, 2), (2, 64, 6, 10)]:
        x = torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype)
        x[0, 0, 0, 0] = float("nan")
        x = x.contiguous(memory_format=torch.channels_last)
        out = torch.empty((B, C, H // 2, W // 2), dtype=dtype, device="cuda", memory_format=torch.channels_last)
        assert L.rcnn_maxpool2x2_nhwc(x.data_ptr(), 1 if dtype == torch.bfloat16 else 0, B, H, W, C, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream) == 0
        want = torch.nn.functional.max_pool2d(x, 2, 2)
        assert torch.equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(want, nan=7.0))
