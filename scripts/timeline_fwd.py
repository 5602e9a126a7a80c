"""Per-timestep timeline of the UNFUSED forward kernel (lstm_fwd_kernel, xp precomputed by K1); the default path is
lstm_fwdx_kernel (scripts/timeline_fwdx.py).  An input width above 512 selects the unfused kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import _lib
B, T, I, H = 256, 64, 576, 512
torch.manual_seed(0)
blk = R.BidirectionalLSTM(I, H, H).cuda()
x = torch.randn(B, T, I, device="cuda")
names = ["P0 start issue", "P1 issued all", "M0 first full", "M1 last commit", "E0 tmem_full", "E1 cell done", "E2 after barrier"]
for mode in ("train", "infer"):
    for it in range(3):
        if mode == "train":
            y = blk(x.requires_grad_(True))
        else:
            with torch.no_grad(): y = blk(x)
    tl = torch.zeros(T * 8 + T * 32, dtype=torch.int64, device="cuda")   # + per-source arrival slots (fwdx)
    _lib.lib().rcnn_debug_timeline(tl.data_ptr())
    if mode == "train":
        y = blk(x.requires_grad_(True))
    else:
        with torch.no_grad(): y = blk(x)
    torch.cuda.synchronize()
    _lib.lib().rcnn_debug_timeline(None)
    a = tl.cpu().numpy().reshape(T, 8).astype(np.float64)
    print(mode, "forward: per-step intervals (cycles), median over steps 2..T-1, relative to P0:")
    for k in range(1, 7):
        d = a[2:, k] - a[2:, 0]
        print(f"  {names[k]:18s} {np.median(d):9.0f}  (min {d.min():.0f} max {d.max():.0f})")
    print("  step period       ", np.median(np.diff(a[2:, 0])))
