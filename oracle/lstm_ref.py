"""Explicit-equation BiLSTM block in torch (CPU, float64 by default) with autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates what
``BidirectionalLSTM.forward`` (model/model.py:159-163) computes through
``nn.LSTM(bidirectional=True, batch_first=True)`` + ``nn.Linear`` as the textbook
cell equations, so that torch autograd over these few lines is the backward oracle
for the BPTT kernels:

    gates = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh      (row blocks i, f, g, o)
    c_t = sigm(f) * c_{t-1} + sigm(i) * tanh(g);  h_t = sigm(o) * tanh(c_t)
    out = [h_fwd ; h_rev] W_lin^T + b_lin,             h_0 = c_0 = 0

Pinned against the reference's own module in tests/golden (tests/make_golden.py).
"""
from __future__ import annotations

import torch


def _direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    xp = x @ w_ih.t() + (b_ih + b_hh)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = xp[:, t] + h @ w_hh.t()
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm_block(x, params: dict, prefix: str = ""):
    """x [B,T,I] -> [B,T,O]; ``params`` holds the reference's state-dict keys."""
    p = lambda k: params[prefix + k]
    hf = _direction(x, p("rnn.weight_ih_l0"), p("rnn.weight_hh_l0"),
                    p("rnn.bias_ih_l0"), p("rnn.bias_hh_l0"), False)
    hr = _direction(x, p("rnn.weight_ih_l0_reverse"), p("rnn.weight_hh_l0_reverse"),
                    p("rnn.bias_ih_l0_reverse"), p("rnn.bias_hh_l0_reverse"), True)
    return torch.cat([hf, hr], dim=2) @ p("linear.weight").t() + p("linear.bias")


def enc_rnn(x, params: dict):
    """RCNN.enc_rnn (model/model.py:195-198): two stacked blocks, keys '0.*' and '1.*'."""
    return bilstm_block(bilstm_block(x, params, "0."), params, "1.")
