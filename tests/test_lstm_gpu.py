"""GPU parity: BidirectionalLSTM block (K1 GEMMs + K2 recurrent kernels) vs the float64
explicit-equation oracle (oracle/lstm_ref.py, pinned to the reference module in
tests/golden).  Tolerance (north_star): 1e-2 absolute on outputs with bf16 operands."""
import numpy as np
import pytest
import torch

from oracle import lstm_ref
from rcnn_ocr_b200 import ops

pytestmark = pytest.mark.gpu

ATOL = 1e-2


def _params(I, H, O, seed, scale=None):
    g = torch.Generator().manual_seed(seed)
    k = scale if scale is not None else 1.0 / np.sqrt(H)
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * k
    p = {}
    for sfx in ("", "_reverse"):
        p["rnn.weight_ih_l0" + sfx] = u(4 * H, I)
        p["rnn.weight_hh_l0" + sfx] = u(4 * H, H)
        p["rnn.bias_ih_l0" + sfx] = u(4 * H)
        p["rnn.bias_hh_l0" + sfx] = u(4 * H)
    p["linear.weight"] = u(O, 2 * H)
    p["linear.bias"] = u(O)
    return p


def _hcat_oracle(x, p):
    xd = x.double()
    pd = {k: v.double() for k, v in p.items()}
    hf = lstm_ref._direction(xd, pd["rnn.weight_ih_l0"], pd["rnn.weight_hh_l0"], pd["rnn.bias_ih_l0"],
                             pd["rnn.bias_hh_l0"], False)
    hr = lstm_ref._direction(xd, pd["rnn.weight_ih_l0_reverse"], pd["rnn.weight_hh_l0_reverse"],
                             pd["rnn.bias_ih_l0_reverse"], pd["rnn.bias_hh_l0_reverse"], True)
    return torch.cat([hf, hr], 2)


@pytest.mark.parametrize("B,T,I,H", [(4, 3, 64, 64), (128, 5, 64, 64), (130, 7, 96, 128), (32, 16, 512, 256),
                                     (256, 64, 512, 512), (1, 1, 64, 64), (3, 33, 128, 512)])
def test_recurrent_forward_matches_oracle(B, T, I, H):
    p = _params(I, H, H, seed=B + T + H)
    x = torch.randn(B, T, I, generator=torch.Generator().manual_seed(1))
    want = _hcat_oracle(x, p)
    pc = {k: v.cuda() for k, v in p.items()}
    packed = ops.lstm_pack(*[pc["rnn." + n + sfx] for sfx in ("", "_reverse")
                             for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")])
    xb = ops.cast_bf16_3d(x.cuda())
    xp = ops.gemm_bf16(xb.view(B * T, I), packed.wih_p, packed.bias_p)
    for save in (False, True):
        hcat, gates, cs = ops.lstm_forward(xp, packed, B, T, save)
        err = (hcat.float().cpu().double() - want).abs().max().item()
        assert err < ATOL, f"max |h - oracle| = {err}"
        if save:
            assert torch.isfinite(gates.float()).all() and torch.isfinite(cs).all()
            # the saved cell state at the last processed step reproduces h = o * tanh(c)
            o = gates[0, T - 1].float().view(B, H // 32, 32, 4)[..., 3].reshape(B, H)
            hrec = o * torch.tanh(cs[0, T - 1])
            assert (hrec - hcat[:, T - 1, :H].float()).abs().max().item() < ATOL
