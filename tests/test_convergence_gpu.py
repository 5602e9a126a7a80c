"""Convergence A/B -- the stand-in for the north_star's "end-to-end val accuracy unchanged" (the reference's validation
loop is training/train.py:531-611; there is no dataset or checkpoint in this environment).

A learnable synthetic line-recognition task is trained twice from the same initial weights, on the same stream of
batches, with the same optimiser settings (Adam, configs/config.json:26-29 scaled up for a short run):
  A. this build: enc_rnn + CTC head + fused CTC loss through the sm_100a kernels (bf16 operands, fp32 state);
  B. the reference's op sequence in fp32 torch (oracle/ref_port: nn.LSTM + nn.Linear x2, log_softmax + F.ctc_loss),
     run on the same GPU so that 300 steps take seconds.
Asserted: both converge, the final training loss (mean of the last 20 steps) agrees to 5 % or 0.02 absolute, and the
validation CER / sequence accuracy of the two models on held-out batches (greedy decode through K4, metrics through
K5) agree to 0.01 absolute.  The measured values are printed (pytest -s / the junit log)."""
import numpy as np
import pytest
import torch

import rcnn_ocr_b200 as R
from oracle import ref_port

pytestmark = pytest.mark.gpu

C, IN, H, T, B = 24, 64, 64, 32, 64          # classes incl. blank, feature size, hidden, frames, batch
STEPS = 300


def _batch(seed, emb):
    """Lines of 3-8 characters; a character occupies a run of frames whose features are its embedding + noise, runs are
    separated by a blank frame pattern (zeros + noise).  Returns feats [B,T,IN], padded targets, lengths."""
    g = torch.Generator().manual_seed(seed)
    tl = torch.randint(3, 9, (B,), generator=g)
    tg = torch.randint(1, C, (B, 8), generator=g)
    feats = torch.randn(B, T, IN, generator=g) * 0.5
    for b in range(B):
        L = int(tl[b])
        width = T // L
        for k in range(L):
            lo = k * width + 1                       # first frame of the run stays "blank"
            feats[b, lo:(k + 1) * width] += emb[tg[b, k]]
    il = torch.full((B,), T)
    return feats, tg, il, tl


def _validate(enc, head, emb, table):
    enc.eval(); head.eval()
    cers, accs = [], []
    with torch.no_grad():
        for seed in range(9000, 9004):
            feats, tg, il, tl = _batch(seed, emb)
            logits = head(enc(feats.cuda())).float()
            ids, lens = R.ctc_greedy_ids(logits)
            m = R.validation_metrics(ids, lens, tg.cuda(), tl, table)
            cers.append(m["cer"]); accs.append(m["accuracy"])
    enc.train(); head.train()
    return float(np.mean(cers)), float(np.mean(accs))


def test_training_converges_like_the_fp32_reference_path():
    torch.manual_seed(0)
    emb = torch.randn(C, IN, generator=torch.Generator().manual_seed(7)) * 1.5
    ours_enc = R.make_enc_rnn(IN, H).cuda()
    ours_head = R.CTCHead(H, C).cuda()
    ref_enc = ref_port.make_encoder(IN, H).cuda()
    ref_head = torch.nn.Linear(H, C).cuda()
    ref_enc.load_state_dict(ours_enc.state_dict(), strict=True)      # same names, same initial weights
    ref_head.load_state_dict(ours_head.state_dict(), strict=True)     # CTCHead registers `weight` / `bias` like nn.Linear
    po = list(ours_enc.parameters()) + list(ours_head.parameters())
    pr = list(ref_enc.parameters()) + list(ref_head.parameters())
    oo = torch.optim.Adam(po, lr=2e-3, weight_decay=1.95e-5)
    orf = torch.optim.Adam(pr, lr=2e-3, weight_decay=1.95e-5)
    hist_o, hist_r = [], []
    for it in range(STEPS):
        feats, tg, il, tl = _batch(100 + it, emb)
        fc, tc = feats.cuda(), tg.cuda()
        oo.zero_grad(set_to_none=True)
        loss = R.ctc_loss_from_logits(ours_head(ours_enc(fc)).permute(1, 0, 2), tc, il, tl, 0, "mean", True)
        loss.backward()
        oo.step()
        hist_o.append(loss.item())
        hist_r.append(ref_port.train_step(ref_enc, ref_head, fc, tc, il, tl))
        orf.step()
    table = R.CharsetTable([chr(0x61 + i) for i in range(C - 1)], torch.device("cuda"))
    fin_o, fin_r = float(np.mean(hist_o[-20:])), float(np.mean(hist_r[-20:]))
    cer_o, acc_o = _validate(ours_enc, ours_head, emb, table)
    cer_r, acc_r = _validate(ref_enc, ref_head, emb, table)
    print(f"\nconvergence A/B over {STEPS} steps: first loss ours {hist_o[0]:.4f} / fp32 {hist_r[0]:.4f}; "
          f"final (last 20) ours {fin_o:.4f} / fp32 {fin_r:.4f}; val CER ours {cer_o:.4f} / fp32 {cer_r:.4f}; "
          f"val accuracy ours {acc_o:.4f} / fp32 {acc_r:.4f}")
    assert abs(hist_o[0] - hist_r[0]) <= 1e-2 * hist_r[0], (hist_o[0], hist_r[0])     # same start (bf16 forward)
    assert fin_r < 0.2 * hist_r[0] and fin_o < 0.2 * hist_o[0], "the task must be learnt by both"
    assert abs(fin_o - fin_r) <= max(0.05 * fin_r, 0.02), (fin_o, fin_r)
    assert abs(cer_o - cer_r) <= 0.01 and abs(acc_o - acc_r) <= 0.02, (cer_o, cer_r, acc_o, acc_r)
