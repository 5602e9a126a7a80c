"""Kernel-level trace (CUPTI) of one greedy attention decode (B = 256, T_enc = 64, H = 512, V = 194, 26 steps) replayed as a CUDA graph."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import rcnn_ocr_b200 as R
torch.manual_seed(1)
B, T, H, V = 256, 64, 512, 194
attn = R.Attention(H, H, V, 1, 2, 0, 3).cuda().eval()
enc = torch.randn(B, T, H, device="cuda")
dec = lambda e: attn(e, is_train=False, batch_max_length=25)
g = R.GraphedStep(dec, [enc])
for _ in range(3): g(enc)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(enc); torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start], key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in evs:
    n = e.name.replace("(anonymous namespace)::", "").replace("rcnn::", "").replace("void ", "").split("(")[0][:50]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span:.1f} us, busy {sum(a[1] for a in agg.values()):.1f} us, {len(evs)} activities")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{a[0]:4d} {a[1]:9.1f} us  {a[1]/a[0]:7.2f} us each  {n}")
t0 = evs[0].time_range.start
for e in evs[10:24]: print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name.split('(')[0][-40:]}")
