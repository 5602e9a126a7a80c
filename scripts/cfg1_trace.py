"""Kernel-level trace (CUPTI) of one cfg-1 inference batch (RCNN(194, 256), 32 lines of 32 x 128, folded bf16 backbone, graph replay)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import rcnn_ocr_b200 as R
torch.manual_seed(0)
model = R.RCNN(194, hidden_size=256).cuda().eval().to(memory_format=torch.channels_last)
model.fold_backbone(torch.bfloat16)
x = (torch.rand(32, 3, 32, 128, device="cuda") * 2 - 1).contiguous(memory_format=torch.channels_last)
@torch.no_grad()
def fwd(x):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        feats = model._features(x)
    return R.ctc_greedy_ids(model.ctc_head(model._encode_features(feats)))
g = R.GraphedStep(fwd, [x])
for _ in range(3): g(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(x); torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start], key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in evs:
    n = e.name.replace("(anonymous namespace)::", "").replace("rcnn::", "").replace("void ", "").split("(")[0][:70]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"span {span:.1f} us, busy {sum(a[1] for a in agg.values()):.1f} us, {len(evs)} activities")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]: print(f"{a[0]:4d} {a[1]:9.1f} us  {a[1]/a[0]:7.2f} us each  {n}")
