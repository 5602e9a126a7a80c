"""Timed CPU baseline: the reference's own torch-CPU op sequence for the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); used by bench.py's ``cpu_baseline``
leg and ``--impl reference`` arm.  /root/reference is Python and cannot travel to the
GPU box, so this is a *port* (kind="port"): the same library calls the reference makes
on CPU, nothing hand-optimised:

  encoder  model/model.py:151-163,195-198   nn.LSTM(bidirectional, batch_first) + nn.Linear, x2
  loss     the north_star's nn.CTCLoss call site replacing training/train.py:289,503-505
           (log_softmax + F.ctc_loss, torch CPU)
  decode   training/utils.py:122-150         argmax + the per-element Python loop
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefBlock(nn.Module):
    """Parameter names match the reference block so state dicts interchange."""

    def __init__(self, n_in: int, n_hidden: int, n_out: int):
        super().__init__()
        self.rnn = nn.LSTM(n_in, n_hidden, bidirectional=True, batch_first=True)
        self.linear = nn.Linear(2 * n_hidden, n_out)

    def forward(self, x):
        seq, _ = self.rnn(x)
        return self.linear(seq)


def make_encoder(n_in: int, hidden: int) -> nn.Sequential:
    return nn.Sequential(RefBlock(n_in, hidden, hidden), RefBlock(hidden, hidden, hidden))


def greedy_decode_loop(logits: torch.Tensor, alphabet, blank: int = 0):
    """The reference's host loop (one .item() per frame)."""
    if logits.dim() == 3 and logits.shape[0] < logits.shape[1]:
        logits = logits.permute(1, 0, 2)
    best = logits.argmax(dim=2)
    texts, seqs = [], []
    for row in best:
        last, seq = blank, []
        for v in row:
            k = v.item()
            if k != blank and k != last:
                seq.append(k)
            last = k
        seqs.append(seq)
        texts.append("".join(alphabet[k - 1] for k in seq))
    return texts, seqs


def train_step(encoder, head, feats, targets, in_len, tg_len, blank=0):
    """fwd + CTC + bwd on CPU; returns the loss value."""
    for p in list(encoder.parameters()) + list(head.parameters()):
        p.grad = None
    logits = head(encoder(feats))                       # [B,T,C]
    lp = F.log_softmax(logits, dim=2).permute(1, 0, 2)  # [T,B,C]
    loss = F.ctc_loss(lp, targets, in_len, tg_len, blank=blank, reduction="mean",
                      zero_infinity=True)
    loss.backward()
    return float(loss)


@torch.no_grad()
def infer_step(encoder, head, feats, alphabet, blank=0):
    logits = head(encoder(feats))
    return greedy_decode_loop(logits, alphabet, blank)


class RefAttention(nn.Module):
    """The reference's attention decoder restated as its torch op sequence (model/model.py:23-108, eval
    mode, greedy path), parameter names as in the reference so state dicts interchange.  Pinned against the
    reference module's own outputs (tests/golden/attn_*.npz, tests/test_oracle_golden.py)."""

    def __init__(self, input_size, hidden_size, num_classes, sos_id=1, blank_id=3):
        super().__init__()
        cell = nn.Module()
        cell.i2h = nn.Linear(input_size, hidden_size, bias=False)
        cell.h2h = nn.Linear(hidden_size, hidden_size)
        cell.score = nn.Linear(hidden_size, 1, bias=False)
        cell.rnn = nn.LSTMCell(input_size + num_classes, hidden_size)
        self.attention_cell = cell
        self.generator = nn.Linear(hidden_size, num_classes)
        self.hidden_size, self.num_classes, self.sos_id, self.blank_id = hidden_size, num_classes, sos_id, blank_id

    @torch.no_grad()
    def greedy(self, batch_H, batch_max_length=25):
        cell = self.attention_cell
        B = batch_H.size(0)
        steps = batch_max_length + 1
        h = batch_H.new_zeros(B, self.hidden_size)
        c = batch_H.new_zeros(B, self.hidden_size)
        y = torch.full((B,), self.sos_id, dtype=torch.long)
        probs = batch_H.new_zeros(B, steps, self.num_classes)
        for t in range(steps):
            onehot = F.one_hot(y, self.num_classes).to(batch_H.dtype)
            proj_H = cell.i2h(batch_H)                                   # recomputed every step, as the reference does
            e = cell.score(torch.tanh(proj_H + cell.h2h(h).unsqueeze(1)))
            alpha = F.softmax(e, dim=1)
            context = torch.bmm(alpha.transpose(1, 2), batch_H).squeeze(1)
            h, c = cell.rnn(torch.cat([context, onehot], 1), (h, c))
            logits = self.generator(h)
            if self.blank_id is not None:
                logits[:, self.blank_id] = -1e4
            probs[:, t] = logits
            y = logits.argmax(1)
        return probs


class RefRCNN(nn.Module):
    """cfg 1 (minimal_inference.py:13-15 -> inference.py:155-180) on the CPU: SE-ResNet31 (the state-dict compatible
    restatement of model/seresnet31.py:70-187, pinned to the reference module in tests/test_oracle_golden.py) ->
    mean over height -> the reference's encoder blocks -> a CTC head (the north_star's decoder) -> the reference's
    greedy loop.  Timed as bench.py's cfg-1 CPU baseline."""

    def __init__(self, num_classes: int, hidden: int):
        super().__init__()
        import rcnn_ocr_b200 as R          # only the pure-torch backbone restatement; no kernels on this path
        self.cnn = R.SEResNet31(3, 512)
        self.enc_rnn = make_encoder(512, hidden)
        self.head = nn.Linear(hidden, num_classes + 1)

    @torch.no_grad()
    def infer(self, x, alphabet):
        feats = self.cnn(x).mean(dim=2).permute(0, 2, 1)
        return greedy_decode_loop(self.head(self.enc_rnn(feats)), alphabet, 0)
