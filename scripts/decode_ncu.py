"""Greedy CTC decode (K4) alone at B = 16384, T = 64, C = 195: a few launches per dtype for ncu / event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rcnn_ocr_b200 as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for dt in (torch.bfloat16, torch.float32):
    xs = [torch.randn(B, 64, 195, device="cuda").to(dt) for _ in range(2)]
    for i in range(3):
        R.ctc_greedy_ids(xs[i % 2])
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); R.ctc_greedy_ids(xs[i % 2]); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    byt = xs[0].numel() * xs[0].element_size()
    print(f"{dt} B={B}: {np.median(ts):.1f} us  {byt / np.median(ts) / 1e3:.0f} GB/s")
