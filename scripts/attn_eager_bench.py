"""Greedy attention decode, eager (no graph): wall time per call with the step loop in C++ (rcnn_attn_greedy_decode) and in Python."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcnn_ocr_b200 as R
torch.manual_seed(1)
attn = R.Attention(512, 512, 194, 1, 2, 0, 3).cuda().eval()
enc = torch.randn(256, 64, 512, device="cuda")
for mode in ("0", "1"):
    os.environ["RCNN_ATTN_PYLOOP"] = mode
    with torch.no_grad():
        for _ in range(5):
            attn(enc, is_train=False, batch_max_length=25)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            attn(enc, is_train=False, batch_max_length=25)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"{'python loop' if mode == '1' else 'C++ loop   '}: {dt * 1e3:.3f} ms per decode of 256 lines, {256 / dt:.0f} lines/s (eager, wall clock)")
