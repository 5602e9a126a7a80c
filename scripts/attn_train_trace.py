"""Kernel-level trace (CUPTI) of one attention-decoder training step (B = 256, T_enc = 64, H = 512, V = 194, 26 steps; teacher forcing,
alpha dropout 0.1, cross entropy, backward) replayed as a CUDA graph."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile
import rcnn_ocr_b200 as R
torch.manual_seed(1)
B, T, H, V, S = 256, 64, 512, 194, 26
attn = R.Attention(H, H, V, 1, 2, 0, 3, dropout_p=0.1).cuda().train()
enc = torch.randn(B, T, H, device="cuda")
text = torch.randint(4, V, (B, S + 1), device="cuda"); text[:, 0] = 1
def step(e, tx):
    attn.zero_grad(set_to_none=True)
    logits = attn(e.detach().requires_grad_(True), tx[:, :S], is_train=True, batch_max_length=S - 1)
    loss = F.cross_entropy(logits.reshape(-1, V), tx[:, 1:].reshape(-1), ignore_index=0)
    loss.backward()
    return loss
g = R.GraphedStep(step, [enc, text])
for _ in range(3): g(enc, text)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(enc, text); torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start], key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in evs:
    n = e.name.replace("(anonymous namespace)::", "").replace("rcnn::", "").replace("void ", "").split("(")[0][:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span:.1f} us, busy {sum(a[1] for a in agg.values()):.1f} us, {len(evs)} activities")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]: print(f"{a[0]:4d} {a[1]:9.1f} us  {a[1]/a[0]:7.2f} us each  {n}")
