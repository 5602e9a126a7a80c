"""GPU parity: attention decoder (K6 + K1 GEMMs) vs the reference's own Attention module
(model/model.py:50-148, imported in the build container by tests/make_golden.py; eval mode, float32).
bf16 GEMM operands: logits within 3e-2 + 2e-2*|ref| of the reference; greedy tokens identical wherever
the reference's own top-2 margin exceeds that error at every earlier step."""
import glob
import os

import numpy as np
import pytest
import torch

import rcnn_ocr_b200 as R
from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu


def _load(name):
    d = golden(name)
    B, T, C, H, V, steps, blank = [int(v) for v in d["dims"]]
    blank = None if blank < 0 else blank
    if "seed_scale" in d.files:
        seed, scale = d["seed_scale"]
        torch.manual_seed(int(seed))
        m = R.Attention(C, H, V, 1, 2, 0, blank, dropout_p=0.1)
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(float(scale))
    else:
        m = R.Attention(C, H, V, 1, 2, 0, blank, dropout_p=0.1)
        sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
        m.load_state_dict(sd, strict=True)            # the reference's keys and shapes, as in a checkpoint
    return d, m.cuda().eval(), steps


def _names():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "attn_*.npz"))
                  if not os.path.basename(p).startswith("attn_train_"))


@pytest.mark.parametrize("name", _names())
def test_teacher_forced_logits_match_reference(name):
    d, m, steps = _load(name)
    x, text = torch.from_numpy(d["batch_H"]).cuda(), torch.from_numpy(d["text"]).cuda()
    want = d["logits"]
    with torch.no_grad():                      # the fused step kernels (K6), as the validation loop runs it (train.py:546-559)
        got = m(x, text=text, is_train=True, batch_max_length=steps - 1)
    assert not got.requires_grad
    graph = m(x, text=text, is_train=True, batch_max_length=steps - 1)     # autograd enabled: the recorded path
    assert graph.requires_grad
    for out in (got, graph.detach()):
        out = out.cpu().numpy()
        assert out.shape == want.shape
        err = np.abs(out - want)
        assert (err <= 3e-2 + 2e-2 * np.abs(want)).all(), f"max |diff| {err.max():.4f}"


@pytest.mark.parametrize("name", _names())
def test_greedy_decode_matches_reference(name):
    """The greedy path = the teacher-forced path fed with its own argmax.  (1) Fed with the REFERENCE's greedy
    tokens our decoder reproduces the reference's greedy probs at every step (no tie sensitivity);
    (2) the feedback argmax is bit-exact (torch.argmax semantics, blank masked first); (3) the free-running
    greedy decode agrees with the reference wherever the reference's top-2 margin stayed above the bf16 error."""
    d, m, steps = _load(name)
    x = torch.from_numpy(d["batch_H"]).cuda()
    want = d["probs"]
    ref_tok = torch.from_numpy(want.argmax(2))
    text = torch.cat([torch.full((want.shape[0], 1), m.sos_id, dtype=torch.int64), ref_tok[:, :-1]], 1)
    with torch.no_grad():
        forced = m(x, text=text.cuda(), is_train=True, batch_max_length=steps - 1).cpu().numpy()
    err = np.abs(forced - want)
    assert (err <= 3e-2 + 2e-2 * np.abs(want)).all(), f"max |diff| {err.max():.4f}"

    got = m(x, is_train=False, batch_max_length=steps - 1)
    assert got.shape == want.shape
    if m.blank_id is not None:
        assert (got[:, :, m.blank_id] == -1e4).all()
    # (2) every fed-back token is the argmax of the probs row that was written for that step
    own_tok = got.argmax(2)
    with torch.no_grad():
        again = m(x, text=torch.cat([torch.full_like(own_tok[:, :1], m.sos_id), own_tok[:, :-1]], 1), is_train=True,
                  batch_max_length=steps - 1)
    assert (again - got).abs().max().item() < 1e-3        # same kernels, same inputs: the greedy run fed back own_tok
    # (3) agreement with the reference while its margins are clear of the bf16 error
    top2 = np.sort(want, axis=2)[:, :, -2:]
    safe = np.minimum.accumulate((top2[:, :, 1] - top2[:, :, 0]) > 0.15, axis=1)
    assert (own_tok.cpu().numpy()[safe] == want.argmax(2)[safe]).all()


def test_argmax_kernel_is_bit_exact():
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    for B, V, blank in [(1, 1, -1), (7, 20, 3), (300, 194, 3), (64, 195, 0), (33, 1000, 999)]:
        x = torch.randint(-3, 4, (B, V), device="cuda", generator=g).float()       # many ties
        x[B // 2, V // 2] = float("nan") if V > 2 else x[B // 2, V // 2]
        y = torch.empty((B,), dtype=torch.int64, device="cuda")
        probs = torch.empty((B, V), device="cuda")
        rc = L.rcnn_attn_argmax(x.data_ptr(), B, V, blank, probs.data_ptr(), V, y.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        want = x.clone()
        if blank >= 0:
            want[:, blank] = -1e4
        assert torch.equal(torch.nan_to_num(probs, nan=7.0), torch.nan_to_num(want, nan=7.0))
        assert torch.equal(y, want.argmax(1))


@pytest.mark.parametrize("B,T,H,C", [(5, 13, 64, 72), (3, 70, 264, 520), (256, 64, 256, 256), (2, 1, 8, 8),
                                     (2, 150, 64, 512)])        # the last: too large to stage in shared memory
def test_score_context_kernels_match_the_formula(B, T, H, C):
    """K6a (both operand widths) against model/model.py:35-41 in float64 on the same (already rounded) inputs:
    alpha to 2e-3 (tanh.approx / __expf), context to bf16 rounding of the float64 value."""
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    projH = torch.randn(B, T, H, device="cuda", generator=g).bfloat16()
    enc = torch.randn(B, T, C, device="cuda", generator=g).bfloat16()
    projh = torch.randn(B, H + 24, device="cuda", generator=g)[:, 8:8 + H]          # a column block of a wider row
    v = torch.randn(H, device="cuda", generator=g) / H ** 0.5
    e = (torch.tanh(projH.double() + projh.double().unsqueeze(1)) * v.double()).sum(-1)
    alpha = torch.softmax(e, 1)
    ctx = (alpha.unsqueeze(2) * enc.double()).sum(1)
    s = torch.cuda.current_stream().cuda_stream
    for half in (True, False):
        a_out = torch.empty(B, T, device="cuda")
        xcat = torch.zeros(B, C + H, dtype=torch.bfloat16, device="cuda")
        if half:
            rc = L.rcnn_attn_score_context_bf16(projH.data_ptr(), projh.data_ptr(), projh.stride(0), v.data_ptr(), enc.data_ptr(),
                                                enc.stride(0), enc.stride(1), B, T, H, C, a_out.data_ptr(), xcat.data_ptr(),
                                                xcat.stride(0), s)
        else:
            pf, ef = projH.float(), enc.float()
            rc = L.rcnn_attn_score_context_ld(pf.data_ptr(), projh.data_ptr(), projh.stride(0), v.data_ptr(), ef.data_ptr(),
                                              ef.stride(0), ef.stride(1), B, T, H, C, a_out.data_ptr(), xcat.data_ptr(),
                                              xcat.stride(0), s)
        assert rc == 0
        assert (a_out.double() - alpha).abs().max().item() <= 2e-3, half
        got = xcat[:, :C].double()
        assert ((got - ctx).abs() <= 2e-3 + 2 ** -7 * ctx.abs()).all(), (half, (got - ctx).abs().max().item())
        assert (xcat[:, C:] == 0).all()                               # the h half of the row is not touched
    # the step entry: the same kernel with the previous step's mask + copy + argmax riding along (bit-exact, K6c's rule)
    V, blank = 37, 3
    lg = torch.randint(-3, 4, (B, V + 11), device="cuda", generator=g).float()[:, 5:5 + V]
    probs = torch.empty(B, 4, V, device="cuda")
    y = torch.full((B,), -1, dtype=torch.int64, device="cuda")
    x2 = torch.zeros_like(xcat)
    xh = torch.zeros_like(xcat)
    assert L.rcnn_attn_score_context_bf16(projH.data_ptr(), projh.data_ptr(), projh.stride(0), v.data_ptr(), enc.data_ptr(),
                                          enc.stride(0), enc.stride(1), B, T, H, C, None, xh.data_ptr(), xh.stride(0), s) == 0
    assert L.rcnn_attn_step_bf16(projH.data_ptr(), projh.data_ptr(), projh.stride(0), v.data_ptr(), enc.data_ptr(), enc.stride(0),
                                 enc.stride(1), B, T, H, C, None, x2.data_ptr(), x2.stride(0), lg.data_ptr(), lg.stride(0), V,
                                 blank, probs[:, 2].data_ptr(), probs.stride(0), y.data_ptr(), s) == 0
    masked = lg.clone()
    masked[:, blank] = -1e4
    assert torch.equal(probs[:, 2], masked) and torch.equal(y, masked.argmax(1)) and torch.equal(x2, xh)
    # shapes without 16-byte rows are refused, not run slowly
    bad = L.rcnn_attn_score_context_bf16(projH.data_ptr(), projh.data_ptr(), projh.stride(0), v.data_ptr(), enc.data_ptr(),
                                         enc.stride(0), enc.stride(1), B, T, H - 1, C, None, xcat.data_ptr(), xcat.stride(0), s)
    assert bad != 0


@pytest.mark.parametrize("B,H,C,V", [(256, 512, 512, 194), (5, 8, 16, 3), (130, 72, 40, 50)])
def test_gate_product_with_cell_epilogue_equals_the_two_launches(B, H, C, V):
    """rcnn_attn_gates_cell (tcgen05 GEMM over gate-interleaved weights, LSTMCell step in its epilogue) against
    rcnn_gemm_bf16 + rcnn_attn_cell on the same inputs: c, h (bf16) and the f32 copy of h bit for bit."""
    from rcnn_ocr_b200 import ops
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(B + H)
    K = C + H
    xcat = torch.randn(B, K, device="cuda", generator=g).bfloat16()
    wcat = (torch.randn(4 * H, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bcat = torch.randn(4 * H, device="cuda", generator=g)
    embT = torch.randn(V, 4 * H, device="cuda", generator=g)
    y = torch.randint(-1, V + 1, (B,), device="cuda", generator=g)               # out-of-range tokens are clamped
    c0 = torch.randn(B, H, device="cuda", generator=g)
    s = torch.cuda.current_stream().cuda_stream
    # two launches
    gates = ops.gemm_bf16(xcat, wcat, bcat, torch.float32)
    c_a, x_a, hid_a = c0.clone(), xcat.clone(), torch.zeros(B, 3, H, device="cuda")
    assert L.rcnn_attn_cell(gates.data_ptr(), embT.data_ptr(), y.data_ptr(), B, H, V, c_a.data_ptr(), x_a.data_ptr(), x_a.stride(0),
                            C, hid_a[:, 1].data_ptr(), hid_a.stride(0), s) == 0
    # one launch
    w_il = wcat.view(4, H, K).permute(1, 0, 2).reshape(4 * H, K).contiguous()
    b_il = bcat.view(4, H).t().reshape(-1).contiguous()
    e_il = embT.view(V, 4, H).permute(0, 2, 1).reshape(V, 4 * H).contiguous()
    c_b, x_b, hid_b = c0.clone(), torch.zeros_like(xcat), torch.zeros(B, 3, H, device="cuda")
    assert L.rcnn_attn_gates_cell(xcat.data_ptr(), xcat.stride(0), w_il.data_ptr(), w_il.stride(0), b_il.data_ptr(), e_il.data_ptr(),
                                  y.data_ptr(), B, H, K, V, c_b.data_ptr(), x_b[:, C:].data_ptr(), x_b.stride(0),
                                  hid_b[:, 1].data_ptr(), hid_b.stride(0), s) == 0
    assert torch.equal(c_a, c_b) and torch.equal(x_a[:, C:], x_b[:, C:]) and torch.equal(hid_a, hid_b)
    assert (x_b[:, :C] == 0).all()
    # writing h into the operand it is computed from is refused
    assert L.rcnn_attn_gates_cell(xcat.data_ptr(), xcat.stride(0), w_il.data_ptr(), w_il.stride(0), b_il.data_ptr(), e_il.data_ptr(),
                                  y.data_ptr(), B, H, K, V, c_b.data_ptr(), xcat.data_ptr(), xcat.stride(0), None, 0, s) != 0


def test_chained_launches_reproduce_the_serialised_decode_bit_for_bit(monkeypatch):
    """The greedy step loop is launched as a programmatic-dependent chain (each kernel may start while its predecessor drains
    and waits on the device before touching its output).  The kernels hold no atomics, so the chained run must equal the
    fully serialised one bit for bit, every time: 30 repeats at a shape with more CTAs than SMs, and through a graph."""
    torch.manual_seed(11)
    m = R.Attention(512, 512, 194, 1, 2, 0, 3).cuda().eval()
    x = torch.randn(300, 40, 512, device="cuda")
    monkeypatch.setenv("RCNN_ATTN_CHAIN", "0")
    with torch.no_grad():
        want = m(x, is_train=False, batch_max_length=12).clone()
        monkeypatch.setenv("RCNN_ATTN_CHAIN", "1")
        for _ in range(30):
            assert torch.equal(m(x, is_train=False, batch_max_length=12), want)
        g = R.GraphedStep(lambda e: m(e, is_train=False, batch_max_length=12), [x])
        for _ in range(10):
            assert torch.equal(g(x), want)
        monkeypatch.setenv("RCNN_ATTN_PYLOOP", "1")            # the same launches issued from Python instead of rcnn_attn_greedy_decode
        assert torch.equal(m(x, is_train=False, batch_max_length=12), want)
        monkeypatch.delenv("RCNN_ATTN_PYLOOP")
        text = want.argmax(2)
        text = torch.cat([torch.ones_like(text[:, :1]), text[:, :-1]], 1)
        tf_want = None
        for mode in ("0", "1", "1"):
            monkeypatch.setenv("RCNN_ATTN_CHAIN", mode)
            got = m(x, text=text, is_train=True, batch_max_length=12)
            tf_want = got.clone() if tf_want is None else tf_want
            assert torch.equal(got, tf_want)


@pytest.mark.parametrize("B,T,H,C,S", [(3, 7, 16, 24, 2), (130, 33, 136, 72, 3)])
def test_attention_backward_kernels_match_autograd_of_the_formula(B, T, H, C, S):
    """K6e (rcnn_attn_step_bwd) and K6f (rcnn_attn_dprojH) against torch autograd of model/model.py:35-41 in float64 on the same
    (bf16-rounded) inputs: de, d proj_h, dv and d proj_H for S independent steps that share proj_H, with a dropout multiplier."""
    L = R.lib()
    g = torch.Generator(device="cuda").manual_seed(B + T)
    projH = torch.randn(B, T, H, device="cuda", generator=g).bfloat16()
    enc = torch.randn(B, T, C, device="cuda", generator=g).bfloat16()
    ph_all = torch.randn(S, B, H, device="cuda", generator=g)
    v = torch.randn(H, device="cuda", generator=g) / H ** 0.5
    scale = (torch.rand(S, B, T, device="cuda", generator=g) >= 0.25).float() / 0.75
    dctx = torch.randn(S, B, C, device="cuda", generator=g)
    # reference: autograd in float64
    pH = projH.double().requires_grad_(True)
    ph = ph_all.double().requires_grad_(True)
    vv = v.double().requires_grad_(True)
    e = (torch.tanh(pH.unsqueeze(0) + ph.unsqueeze(2)) * vv).sum(-1)                   # [S,B,T]
    e.retain_grad()
    alpha = torch.softmax(e, 2)
    ctx = ((alpha * scale.double()).unsqueeze(3) * enc.double().unsqueeze(0)).sum(2)    # [S,B,C]
    (ctx * dctx.double()).sum().backward()
    # kernels
    s = torch.cuda.current_stream().cuda_stream
    de_all = torch.empty(S, B, T, device="cuda")
    dproj = torch.zeros(S, B, H, dtype=torch.bfloat16, device="cuda")
    dv_acc = torch.zeros(B, H, device="cuda")
    a32 = alpha.detach().float().contiguous()
    for t in range(S):
        assert L.rcnn_attn_step_bwd(dctx[t].data_ptr(), C, a32[t].data_ptr(), scale[t].data_ptr(), enc.data_ptr(), enc.stride(0),
                                    enc.stride(1), projH.data_ptr(), ph_all[t].data_ptr(), H, v.data_ptr(), B, T, H, C,
                                    de_all[t].data_ptr(), dproj[t].data_ptr(), H, dv_acc.data_ptr(), s) == 0
    dprojH = torch.empty(B, T, H, dtype=torch.bfloat16, device="cuda")
    assert L.rcnn_attn_dprojH(de_all.data_ptr(), ph_all.data_ptr(), projH.data_ptr(), v.data_ptr(), S, B, T, H, dprojH.data_ptr(), s) == 0

    def close(got, want, what, rel):
        err = (got.double() - want).abs().max().item()
        assert err <= rel * want.abs().max().item() + 1e-6, (what, err, want.abs().max().item())

    close(de_all, e.grad, "de", 4e-3)
    close(dproj, ph.grad, "d proj_h", 1e-2)              # bf16 output
    close(dv_acc.sum(0), vv.grad, "dv", 4e-3)
    close(dprojH, pH.grad, "d proj_H", 1e-2)             # bf16 output


def test_cell_backward_kernel_matches_autograd():
    """K6d (rcnn_attn_cell_bwd) against autograd of the LSTMCell pointwise step in float64: gate-interleaved pre-activation
    gradients (bf16) and dc_{t-1}, with both dh addends and an incoming dc."""
    L = R.lib()
    B, H = 37, 24
    g = torch.Generator(device="cuda").manual_seed(5)
    pre = torch.randn(B, H, 4, device="cuda", generator=g).double().requires_grad_(True)      # [.., (i, f, g, o)] interleaved
    c_prev = torch.randn(B, H, device="cuda", generator=g).double().requires_grad_(True)
    i, f, gg, o = torch.sigmoid(pre[..., 0]), torch.sigmoid(pre[..., 1]), torch.tanh(pre[..., 2]), torch.sigmoid(pre[..., 3])
    c_t = f * c_prev + i * gg
    h = o * torch.tanh(c_t)
    dh_a = torch.randn(B, 3, H, device="cuda", generator=g)
    dh_b = torch.randn(B, H, device="cuda", generator=g)
    dc_in = torch.randn(B, H, device="cuda", generator=g)
    (h * (dh_a[:, 1] + dh_b).double() + c_t * dc_in.double()).sum().backward()
    acts = torch.stack([i, f, gg, o], -1).detach().float().reshape(B, 4 * H).contiguous()
    dc = dc_in.clone()
    dg = torch.zeros(B, 5 * H, dtype=torch.bfloat16, device="cuda")
    assert L.rcnn_attn_cell_bwd(acts.data_ptr(), c_prev.detach().float().contiguous().data_ptr(), c_t.detach().float().contiguous().data_ptr(),
                                dh_a[:, 1].data_ptr(), dh_a.stride(0), dh_b.data_ptr(), H, dc.data_ptr(), B, H, dg.data_ptr(), 5 * H,
                                torch.cuda.current_stream().cuda_stream) == 0
    want = pre.grad.reshape(B, 4 * H)
    assert (dg[:, :4 * H].double() - want).abs().max().item() <= 1e-2 * want.abs().max().item()
    assert (dc.double() - c_prev.grad).abs().max().item() <= 2e-3 * c_prev.grad.abs().max().item()
    assert (dg[:, 4 * H:] == 0).all()


def test_state_dict_contract_and_errors():
    m = R.Attention(64, 64, 20, 1, 2, 0, 3)
    want = {"attention_cell.i2h.weight": (64, 64), "attention_cell.h2h.weight": (64, 64), "attention_cell.h2h.bias": (64,),
            "attention_cell.score.weight": (1, 64), "attention_cell.rnn.weight_ih": (256, 84),
            "attention_cell.rnn.weight_hh": (256, 64), "attention_cell.rnn.bias_ih": (256,),
            "attention_cell.rnn.bias_hh": (256,), "generator.weight": (20, 64), "generator.bias": (20,)}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == want
    m = m.cuda()
    x = torch.randn(2, 5, 64, device="cuda")
    with pytest.raises(NotImplementedError):
        m.train()(x, is_train=False)                      # dropout / training path is out of scope
    m.eval()
    with pytest.raises(AssertionError):
        m(x, text=None, is_train=True)                    # model/model.py:114-116
    with pytest.raises(RuntimeError):
        m(x.cpu(), is_train=False)                        # no CPU fallback
    assert m(x, is_train=False, batch_max_length=3).shape == (2, 4, 20)


@pytest.mark.parametrize("name", ["attn_train_tf", "attn_train_sampled", "attn_train_mixed"])
def test_training_path_matches_the_reference_gradients(name):
    """model/model.py:110-148 in train() mode: teacher-forced logits and the gradient of every parameter and of batch_H
    against the reference module's own autograd (goldens; dropout 0; scheduled sampling 0 / 1 / 0.5 with the reference's
    one CPU draw per step, so the same seed takes the same decisions).  bf16 GEMM operands: logits 2e-2 absolute (their
    magnitude is ~5), gradients max|diff| <= 2e-2 max|grad| per tensor."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T, C, H, V, steps, blank = [int(v) for v in d["dims"]]
    seed, scale, sampling = int(d["seed"]), float(d["scale"]), float(d["sampling"])
    torch.manual_seed(seed)
    m = R.Attention(C, H, V, 1, 2, 0, None if blank < 0 else blank, dropout_p=0.0, sampling_prob=sampling)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(scale)
    m = m.cuda().train()
    x = torch.from_numpy(d["batch_H"]).cuda().requires_grad_(True)
    text = torch.from_numpy(d["text"]).cuda()
    torch.manual_seed(seed + 2)
    logits = m(x, text=text, is_train=True, batch_max_length=steps - 1)
    assert logits.shape == (B, steps, V) and logits.requires_grad
    (logits * torch.from_numpy(d["w"]).cuda()).sum().backward()
    err = np.abs(logits.detach().cpu().numpy() - d["logits"]).max()
    assert err <= 2e-2, f"logits: {err}"
    worst = {}
    for k, p in list(m.named_parameters()) + [("batch_H", x)]:
        want = d["g." + k]
        got = p.grad.detach().cpu().numpy()
        worst[k] = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    k = max(worst, key=worst.get)
    print(f"\n{name}: logits max|diff| {err:.2e}; gradients max|diff|/max|grad| worst {worst[k]:.2e} ({k})")
    assert worst[k] <= 2e-2, worst


@pytest.mark.parametrize("B,T,C,H,V,steps", [(6, 10, 64, 64, 30, 6), (130, 33, 72, 136, 50, 4)])
def test_fused_training_path_equals_the_autograd_path(B, T, C, H, V, steps, monkeypatch):
    """_TeacherForcedFn (fused forward / backward kernels) against the op-by-op autograd path on the same weights, inputs
    and the SAME alpha dropout mask: logits and every gradient to 2e-2 of the tensor's max (bf16 operands on both)."""
    torch.manual_seed(3)
    m = R.Attention(C, H, V, 1, 2, 0, 3, dropout_p=0.3).cuda().train()
    x0 = torch.randn(B, T, C, device="cuda")
    text = torch.randint(0, V, (B, steps + 1), device="cuda")
    text[:, 0] = 1
    w = torch.randn(B, steps, V, device="cuda")
    m._alpha_scale_override = (torch.rand(steps, B, T, device="cuda") >= 0.3).float() / 0.7
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RCNN_ATTN_FUSED_TRAIN", mode)
        m.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        logits = m(x, text=text, is_train=True, batch_max_length=steps - 1)
        (logits * w).sum().backward()
        res[mode] = {"logits": logits.detach(), "batch_H": x.grad.clone(), **{k: p.grad.clone() for k, p in m.named_parameters()}}
    worst = {k: ((res["1"][k] - res["0"][k]).abs().max() / res["0"][k].abs().max().clamp_min(1e-30)).item() for k in res["0"]}
    k = max(worst, key=worst.get)
    print(f"\nfused vs autograd training path: worst max|diff|/max {worst[k]:.2e} ({k})")
    assert worst[k] <= 2e-2, worst


def test_dropout_and_attention_rcnn_train_step(tmp_path):
    """Dropout is live in train() mode (two passes differ, eval passes agree), and RCNN(decoder="attention") trains the
    way training/train.py:499-505 does: teacher forcing, cross entropy with ignore_index=<PAD>, backward through the
    attention decoder, the encoder blocks and into the feature columns."""
    torch.manual_seed(0)
    m = R.Attention(64, 64, 30, 1, 2, 0, 3, dropout_p=0.5).cuda().train()
    x = torch.randn(4, 10, 64, device="cuda")
    text = torch.randint(4, 30, (4, 6), device="cuda")
    text[:, 0] = 1
    a, b = m(x, text=text, batch_max_length=5), m(x, text=text, batch_max_length=5)
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    m.eval()
    with torch.no_grad():
        c, e = m(x, text=text, batch_max_length=5), m(x, text=text, batch_max_length=5)
    assert torch.equal(c, e)
    # RCNN with the reference's decoder, a few optimiser steps on one batch
    itos = ["<PAD>", "<SOS>", "<EOS>", "<BLANK>"] + [chr(0x61 + i) for i in range(12)]
    stoi = {s: i for i, s in enumerate(itos)}
    model = R.RCNN(num_classes=len(itos), hidden_size=64, decoder="attention").cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    imgs = torch.rand(6, 3, 32, 64, device="cuda") * 2 - 1
    text_in, target_y, _ = R.pack_attention_targets(["abc", "ca", "b", "abba", "lk", "hgf"], stoi, max_len=6)
    crit = torch.nn.CrossEntropyLoss(ignore_index=stoi["<PAD>"])
    hist = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        logits = model(imgs, text=text_in.cuda(), is_train=True, batch_max_length=6)
        loss = crit(logits.reshape(-1, logits.shape[-1]), target_y.cuda().reshape(-1))
        loss.backward()
        opt.step()
        hist.append(loss.item())
    assert all(p.grad is not None for p in model.attn.parameters()) and hist[-1] < hist[0], hist
