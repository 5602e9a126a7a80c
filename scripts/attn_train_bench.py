"""Attention decoder, one teacher-forced training step (forward + cross-entropy + backward), B = 256, T_enc = 64, H = 512,
V = 194, 26 steps: this build's autograd path (K1 GEMMs + torch elementwise) against the reference's op sequence in torch
fp32 on the same GPU (nn.Linear / nn.LSTMCell / bmm, proj_H recomputed every step as model/model.py:35 does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import rcnn_ocr_b200 as R
from oracle.ref_port import RefAttention

B, T, H, V, steps = 256, 64, 512, 194, 26
torch.manual_seed(0)
ours = R.Attention(H, H, V, 1, 2, 0, 3, dropout_p=0.0).cuda().train()
ref = RefAttention(H, H, V).cuda().train()
ref.load_state_dict(ours.state_dict(), strict=True)
enc = torch.randn(B, T, H, device="cuda", requires_grad=True)
text = torch.randint(4, V, (B, steps + 1), device="cuda")
text[:, 0] = 1
tgt = torch.randint(4, V, (B, steps), device="cuda")


def step_ours():
    logits = ours(enc, text[:, :steps], is_train=True, batch_max_length=steps - 1)
    loss = F.cross_entropy(logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    return loss


def step_ours_on(e, tx):
    ours.zero_grad(set_to_none=True)
    logits = ours(e.detach().requires_grad_(True), tx[:, :steps], is_train=True, batch_max_length=steps - 1)
    loss = F.cross_entropy(logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    return loss


def step_ref():
    cell = ref.attention_cell
    h = enc.new_zeros(B, H); c = enc.new_zeros(B, H)
    hs = []
    for t in range(steps):
        onehot = F.one_hot(text[:, t], V).float()
        e = cell.score(torch.tanh(cell.i2h(enc) + cell.h2h(h).unsqueeze(1)))
        alpha = F.softmax(e, dim=1)
        context = torch.bmm(alpha.transpose(1, 2), enc).squeeze(1)
        h, c = cell.rnn(torch.cat([context, onehot], 1), (h, c))
        hs.append(h)
    logits = ref.generator(torch.stack(hs, 1))
    logits[:, :, 3] = -1e4
    loss = F.cross_entropy(logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    return loss


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3


for name, fn in (("ours, eager launches", step_ours),) + ((("torch fp32 eager", step_ref),) if len(sys.argv) < 2 else ()):
    dev, wall = timed(fn)
    print(f"{name:24s} {dev:8.2f} ms per step (wall {wall:.2f})  {B / dev * 1e3:9.0f} lines/s   loss {fn().item():.4f}")
g = R.GraphedStep(lambda e, tx: step_ours_on(e, tx), [enc.detach(), text])
dev, wall = timed(lambda: g(enc.detach(), text), 30)
print(f"{'ours, one graph replay':24s} {dev:8.3f} ms per step (wall {wall:.2f})  {B / dev * 1e3:9.0f} lines/s")
if len(sys.argv) > 1 and sys.argv[1] == "ours":
    sys.exit(0)
for dt in (torch.bfloat16,):
    def ac():
        with torch.autocast("cuda", dtype=dt):
            return step_ref()
    dev, wall = timed(ac)
    print(f"{'torch autocast bf16':24s} {dev:8.2f} ms per step (wall {wall:.2f})  {B / dev * 1e3:9.0f} lines/s")
