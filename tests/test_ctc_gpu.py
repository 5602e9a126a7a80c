"""GPU parity: fused CTC kernel (K3) vs the float64 oracle and torch-float64 golden vectors.
Tolerance (north_star): 1e-4 relative in float32 -- asserted as rtol=1e-4 on the loss and
|diff| <= 1e-4 * max|grad| (+ rtol 1e-4) on gradients taken at the logits."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
import rcnn_ocr_b200 as R
from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _names(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def _assert_grad(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape
    nan_w, nan_g = np.isnan(want), np.isnan(got)
    np.testing.assert_array_equal(nan_g, nan_w)
    scale = np.nanmax(np.abs(want)) if np.isfinite(want).any() else 1.0
    np.testing.assert_allclose(np.where(nan_w, 0, got), np.where(nan_w, 0, want), rtol=RTOL,
                               atol=RTOL * max(scale, 1e-30))


@pytest.mark.parametrize("name", _names("ctc_"))
def test_golden_torch_f64(name):
    d = golden(name)
    blank = int(d["blank"])
    tg = torch.from_numpy(d["targets"]).cuda()
    il, tl = torch.from_numpy(d["input_lengths"]), torch.from_numpy(d["target_lengths"])
    for zi in (0, 1):
        x = torch.tensor(d["x"], dtype=torch.float32, device="cuda", requires_grad=True)
        nll = R.ctc_loss_from_logits(x, tg, il, tl, blank, "none", bool(zi))
        np.testing.assert_allclose(nll.detach().cpu().numpy(), d[f"nll_zi{zi}"], rtol=RTOL)
        for red in ("mean", "sum"):
            if f"grad_{red}_zi{zi}" not in d.files:
                continue
            x = torch.tensor(d["x"], dtype=torch.float32, device="cuda", requires_grad=True)
            loss = R.ctc_loss_from_logits(x, tg, il, tl, blank, red, bool(zi))
            np.testing.assert_allclose(loss.item(), d[f"loss_{red}_zi{zi}"], rtol=RTOL)
            if np.isfinite(loss.item()):
                loss.backward()
                _assert_grad(x.grad.cpu().numpy(), d[f"grad_{red}_zi{zi}"])
    # concatenated targets
    x = torch.tensor(d["x"], dtype=torch.float32, device="cuda", requires_grad=True)
    loss = R.ctc_loss_from_logits(x, torch.from_numpy(d["targets_concat"]), il, tl, blank, "mean", True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), d["loss_mean_zi1"], rtol=RTOL)
    _assert_grad(x.grad.cpu().numpy(), d["grad_mean_zi1"])


def _random_case(T, N, C, max_l, seed, var_in=False, blank=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, C, generator=g)
    tl = torch.randint(1, max_l + 1, (N,), generator=g)
    labels = torch.tensor([c for c in range(C) if c != blank])
    tg = labels[torch.randint(0, len(labels), (N, max_l), generator=g)]
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g) if var_in else torch.full((N,), T)
    return x, tg, il, tl


@pytest.mark.parametrize("T,N,C,max_l,var_in", [
    (64, 256, 195, 32, False),      # BASELINE config 2
    (64, 256, 195, 32, True),
    (16, 128, 195, 8, False),       # configs/config.json geometry (T = 128/8)
    (1, 3, 4, 1, False), (2, 5, 3, 1, False), (33, 17, 50, 16, True), (128, 8, 20, 60, True),
    (40, 6, 300, 40, False), (200, 3, 30, 90, True),
])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_random_vs_oracle(T, N, C, max_l, var_in, reduction):
    x, tg, il, tl = _random_case(T, N, C, max_l, seed=T * 7 + N, var_in=var_in)
    want_loss, want_grad = oracle.ctc_loss(x.numpy().astype(np.float64), tg.numpy(), il.numpy(),
                                           tl.numpy(), 0, reduction, True, from_logits=True)
    xg = x.cuda().requires_grad_(True)
    loss = R.ctc_loss_from_logits(xg, tg.cuda(), il, tl, 0, reduction, True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), want_loss, rtol=RTOL)
    _assert_grad(xg.grad.cpu().numpy(), want_grad)


def test_batch_first_strided_input_and_grad_layout():
    """The training step feeds head_out[N,T,C].permute(1,0,2) without a copy."""
    x, tg, il, tl = _random_case(24, 10, 37, 9, seed=2, var_in=True)
    want_loss, want_grad = oracle.ctc_loss(x.numpy().astype(np.float64), tg.numpy(), il.numpy(),
                                           tl.numpy(), 0, "mean", True)
    base = x.permute(1, 0, 2).contiguous().cuda().requires_grad_(True)    # [N,T,C] storage
    loss = R.ctc_loss_from_logits(base.permute(1, 0, 2), tg.cuda(), il.cuda(), tl.cuda(), 0, "mean", True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), want_loss, rtol=RTOL)
    _assert_grad(base.grad.permute(1, 0, 2).cpu().numpy(), want_grad)


def test_module_is_a_ctcloss_drop_in():
    """nn.CTCLoss semantics on log-prob inputs, all reductions, vs torch's own module."""
    x, tg, il, tl = _random_case(30, 12, 25, 10, seed=4, var_in=True)
    il[3] = 2; tl[3] = 9                                             # infeasible sample
    for red in ("none", "mean", "sum"):
        for zi in (False, True):
            a = x.double().requires_grad_(True)
            ref = torch.nn.CTCLoss(0, red, zi)(a.log_softmax(2), tg, il, tl)
            b = x.cuda().requires_grad_(True)
            got = R.CTCLoss(0, red, zi)(b.log_softmax(2), tg.cuda(), il, tl)
            np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().numpy(), rtol=RTOL)
            if zi:
                w = torch.linspace(0.5, 2.0, ref.numel()).reshape(ref.shape)
                (ref * w.double()).sum().backward()
                (got * w.cuda()).sum().backward()
                _assert_grad(b.grad.cpu().numpy(), a.grad.numpy())


def test_edge_cases():
    T, N, C = 10, 6, 5
    x = torch.randn(T, N, C, generator=torch.Generator().manual_seed(8))
    tg = torch.tensor([[1, 1, 2, 0], [3, 0, 0, 0], [0, 0, 0, 0], [1, 2, 3, 4], [2, 2, 2, 2], [4, 0, 0, 0]])
    tl = torch.tensor([3, 1, 0, 4, 4, 1])
    il = torch.tensor([10, 1, 10, 4, 7, 0])       # T==L tight, empty target, L+repeats == T, T_n = 0
    for zi in (False, True):
        want_nll, want_g = oracle.ctc_nll_and_grad(x.numpy().astype(np.float64), tg.numpy(), il.numpy(),
                                                   tl.numpy(), 0, zi, True)
        xg = x.cuda().requires_grad_(True)
        nll = R.ctc_loss_from_logits(xg, tg.cuda(), il, tl, 0, "none", zi)
        np.testing.assert_allclose(nll.detach().cpu().numpy(), want_nll, rtol=RTOL)
        nll.sum().backward()
        _assert_grad(xg.grad.cpu().numpy(), want_g)
    # -inf log-probabilities on the only feasible path -> inf loss, handled like torch
    lp = torch.randn(4, 1, 3).log_softmax(2)
    lp[2, 0, 1] = float("-inf")
    ref = torch.nn.functional.ctc_loss(lp.double(), torch.tensor([[1, 1]]), [4], [2], reduction="none")
    got = R.ctc_loss(lp.cuda(), torch.tensor([[1, 1]]).cuda(), [4], [2], reduction="none")
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=RTOL)
    # empty batch
    e = R.ctc_loss_from_logits(torch.zeros(5, 0, 4, device="cuda"), torch.zeros(0, 2, dtype=torch.long),
                               [], [], reduction="sum")
    assert e.item() == 0.0


def test_long_targets_spill_to_workspace():
    x, tg, il, tl = _random_case(300, 4, 40, 120, seed=6, var_in=True)
    want_loss, want_grad = oracle.ctc_loss(x.numpy().astype(np.float64), tg.numpy(), il.numpy(),
                                           tl.numpy(), 0, "sum", True)
    xg = x.cuda().requires_grad_(True)
    loss = R.ctc_loss_from_logits(xg, tg.cuda(), il, tl, 0, "sum", True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), want_loss, rtol=RTOL)
    _assert_grad(xg.grad.cpu().numpy(), want_grad)


def test_full_size_properties():
    """N = 16384 (sweep top): per-frame gradient rows sum to zero inside the input length,
    are zero beyond it, nll is positive, and a slice matches the oracle."""
    T, N, C = 64, 16384, 195
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(T, N, C, device="cuda", generator=g, requires_grad=True)
    tl = torch.randint(1, 33, (N,), device="cuda", generator=g)
    tg = torch.randint(1, C, (N, 32), device="cuda", generator=g)
    il = torch.randint(60, 65, (N,), device="cuda", generator=g)
    nll = R.ctc_loss_from_logits(x, tg, il, tl, 0, "none", True, max_target_length=32)
    nll.sum().backward()
    assert bool((nll > 0).all()) and bool(torch.isfinite(nll).all())
    rows = x.grad.sum(2)
    assert float(rows.abs().max()) < 5e-5
    tmask = torch.arange(T, device="cuda")[:, None] >= il[None, :]
    assert float(x.grad[tmask].abs().max()) == 0.0
    sl = slice(100, 164)
    want_nll, want_g = oracle.ctc_nll_and_grad(x.detach()[:, sl].cpu().numpy().astype(np.float64),
                                               tg[sl].cpu().numpy(), il[sl].cpu().numpy(),
                                               tl[sl].cpu().numpy(), 0, True, True)
    np.testing.assert_allclose(nll[sl].detach().cpu().numpy(), want_nll, rtol=RTOL)
    _assert_grad(x.grad[:, sl].cpu().numpy(), want_g)


def test_low_precision_and_strided_inputs_still_get_gradients():
    """bf16 / fp16 log-probs (autocast) and inputs with a non-unit class stride are converted inside
    forward(); the gradient must still come back (in the input's dtype and shape) and equal the fp32 result."""
    g = torch.Generator().manual_seed(21)
    T, N, C, L = 12, 5, 9, 4
    base = torch.randn(T, N, C, generator=g).cuda()
    tg = torch.randint(1, C, (N, L), generator=g).cuda()
    il, tl = torch.full((N,), T), torch.randint(1, L + 1, (N,), generator=g)

    def run(x):
        x = x.detach().clone().requires_grad_(True)
        loss = R.ctc_loss(x.log_softmax(-1) if x.dtype == torch.float32 else x, tg, il, tl, 0, "sum", True)
        loss.backward()
        return loss.item(), x.grad

    # bf16 log-probs: compare against the fp32 path fed the same (bf16-rounded) values
    lp16 = base.log_softmax(-1).to(torch.bfloat16)
    x16 = lp16.detach().clone().requires_grad_(True)
    loss16 = R.ctc_loss(x16, tg, il, tl, 0, "sum", True)
    loss16.backward()
    assert x16.grad is not None and x16.grad.dtype == torch.bfloat16 and x16.grad.shape == x16.shape
    x32 = lp16.float().detach().clone().requires_grad_(True)
    loss32 = R.ctc_loss(x32, tg, il, tl, 0, "sum", True)
    loss32.backward()
    np.testing.assert_allclose(loss16.item(), loss32.item(), rtol=1e-6)
    np.testing.assert_allclose(x16.grad.float().cpu().numpy(), x32.grad.cpu().numpy(), rtol=1e-2, atol=1e-2)
    assert x16.grad.float().abs().max().item() > 0

    # class stride 2 (a slice of a wider tensor): converted with .contiguous() inside forward()
    wide = torch.zeros(T, N, 2 * C, device="cuda")
    wide[:, :, ::2] = base.log_softmax(-1)
    xs = wide.detach().clone().requires_grad_(True)
    loss_s = R.ctc_loss(xs[:, :, ::2], tg, il, tl, 0, "sum", True)
    loss_s.backward()
    xd = base.log_softmax(-1).detach().clone().requires_grad_(True)
    loss_d = R.ctc_loss(xd, tg, il, tl, 0, "sum", True)
    loss_d.backward()
    np.testing.assert_allclose(loss_s.item(), loss_d.item(), rtol=1e-6)
    np.testing.assert_allclose(xs.grad[:, :, ::2].cpu().numpy(), xd.grad.cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert (xs.grad[:, :, 1::2] == 0).all()


def test_target_overrun_raises_like_torch():
    x = torch.randn(6, 2, 5, device="cuda").log_softmax(-1)
    with pytest.raises(RuntimeError):      # concatenated targets shorter than sum(target_lengths)
        R.ctc_loss(x, torch.tensor([1, 2, 3]), torch.tensor([6, 6]), torch.tensor([2, 2]))
    # padded targets: a length beyond the row width flags that row (NaN), and max_target_length is clamped to the
    # row width, so the kernel never reads past the row
    out = R.ctc_loss(x, torch.tensor([[1, 2], [3, 4]]).cuda(), torch.tensor([6, 6]), torch.tensor([2, 5]), 0, "none", True,
                     max_target_length=5)
    assert torch.isfinite(out[0]) and torch.isnan(out[1])
