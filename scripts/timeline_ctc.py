"""clock64 at the phase boundaries of CTA 0 of the fused CTC kernel (cfg 2: T=64, C=195, labels U{1..32})."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import _lib
T, C = 64, 195
for N in (256, 4096):
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(N, T, C, device="cuda", generator=g).requires_grad_(True) for _ in range(3)]
    tl_ = torch.randint(1, 33, (N,), device="cuda", generator=g)
    tl_[0] = 32
    tg = torch.randint(1, C, (N, 32), device="cuda", generator=g)
    il = torch.full((N,), T, device="cuda")
    def run(i):
        xs[i].grad = None
        R.ctc_loss_from_logits(xs[i].permute(1, 0, 2), tg, il, tl_, 0, "mean", True, max_target_length=32).backward()
    for i in range(3): run(i)
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    _lib.lib().rcnn_debug_timeline(buf.data_ptr())
    run(0)
    torch.cuda.synchronize()
    _lib.lib().rcnn_debug_timeline(None)
    a = buf.cpu().tolist()
    print(f"N={N}: CTA 0 (L=32, S=65): phase 1 {a[1]-a[0]} cycles, phase 2 (alpha/beta) {a[2]-a[1]}, phase 3 {a[3]-a[2]}, total {a[3]-a[0]}")
