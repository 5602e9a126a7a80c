"""Repeatability stress of a whole BidirectionalLSTM block (forward + backward through autograd): out and dx involve
no floating-point atomics and must be bit-identical on every repeat; parameter gradients (split-K reduce-add)
must agree to rounding.  Usage: python scripts/stress_module.py [B T I H O] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rcnn_ocr_b200 as R

B, T, I, H, O = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (2500, 2, 64, 64, 64)
iters = int(sys.argv[6]) if len(sys.argv) >= 7 else 200
torch.manual_seed(0)
blk = R.BidirectionalLSTM(I, H, O).cuda()
x = torch.randn(B, T, I).permute(0, 2, 1).contiguous().cuda().permute(0, 2, 1).requires_grad_(True)
w = torch.randn(B, T, O, device="cuda") / (B * T) ** 0.5
ref = None
bad = 0
for it in range(iters):
    x.grad = None
    blk.zero_grad(set_to_none=True)
    out = blk(x)
    (out * w).sum().backward()
    torch.cuda.synchronize()
    cur = {"out": out.detach(), "dx": x.grad}
    cur.update({k: p.grad for k, p in blk.named_parameters()})
    if ref is None:
        ref = {k: v.clone() for k, v in cur.items()}
        continue
    msgs = []
    for k, v in cur.items():
        d = (v - ref[k]).abs().max().item()
        tol = 0.0 if k in ("out", "dx") else 1e-3 * ref[k].abs().max().item() + 1e-7
        if d > tol:
            msgs.append(f"{k}: {d:.3e} (max {ref[k].abs().max().item():.3e})")
    if msgs:
        bad += 1
        if bad <= 8:
            print(f"iter {it}: " + "; ".join(msgs))
print(f"B={B} T={T} I={I} H={H} O={O}: {bad} of {iters - 1} repeats differ")
