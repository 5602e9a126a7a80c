"""Recurrent kernels alone (K2): CUDA-event time per launch of the fused forward (inference / training variant) and of
the backward kernel over the batch size, with the layout rcnn_lstm_plan picks, plus the clock64 timeline of CTA 0.

    python scripts/k2_bench.py [B ...]            # default 256 512 1024
    RCNN_FWD_SLOTS=1 RCNN_BWD_SLOTS=1 python scripts/k2_bench.py 512     # the one-item-per-group kernels, for A/B
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rcnn_ocr_b200 import _lib, ops

T, I, H = 64, 512, 512
Bs = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(0)
k = 1.0 / H ** 0.5
ws = []
for _ in range(2):
    ws += [(torch.rand(4 * H, I, device=dev, generator=g) * 2 - 1) * k, (torch.rand(4 * H, H, device=dev, generator=g) * 2 - 1) * k,
           (torch.rand(4 * H, device=dev, generator=g) * 2 - 1) * k, (torch.rand(4 * H, device=dev, generator=g) * 2 - 1) * k]
packed = ops.lstm_pack(*ws)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = [e[i].elapsed_time(e[i + 1]) for i in range(reps)]
    return float(np.median(ts)), float(min(ts))


def timeline(fn, names, order):
    tl = torch.zeros(T * 8 + T * 32, dtype=torch.int64, device=dev)
    _lib.lib().rcnn_debug_timeline(tl.data_ptr())
    fn()
    torch.cuda.synchronize()
    _lib.lib().rcnn_debug_timeline(None)
    a = tl.cpu().numpy().astype(np.float64)[:T * 8].reshape(T, 8)[:-1]
    for kk in order:
        d = a[2:, kk] - a[2:, 0]
        print(f"      {names[kk]:44s} {np.median(d):8.0f}  (min {d.min():.0f} max {d.max():.0f})")
    print(f"      step period {np.median(np.diff(a[2:, 0])):.0f} cycles")
    return tl.cpu().numpy().astype(np.float64)


FW = {1: "M0 first h box landed", 2: "M1 half-0 MMAs issued", 3: "M2 half-1 MMAs issued", 4: "E0 half-0 accumulator complete",
      5: "E1 half-0 cell phase done, h stored", 6: "R0 publisher saw h_staged", 7: "R1 release issued"}
BW = {1: "P1 partials stored (wait_group 0)", 2: "M0 dG tile ready: first MMA", 3: "M1 last MMA issued (slot 0)",
      4: "E0 accumulators complete", 5: "E1 cell done (dG stored)", 6: "E2 counter released", 7: "S0 partials staged in smem"}
for B in Bs:
    x = torch.randn(B, T, I, device=dev, generator=g).bfloat16()
    dh = torch.randn(B, T, 2 * H, device=dev, generator=g) / (B * T) ** 0.5
    hcat, gates, cs = ops.lstm_forward_fused(x, packed, B, T, True)
    fi = timed(lambda: ops.lstm_forward_fused(x, packed, B, T, False))
    ft = timed(lambda: ops.lstm_forward_fused(x, packed, B, T, True))
    bw = timed(lambda: ops.lstm_backward(packed, gates, cs, dh, B, T))
    flop = 2.0 * B * T * H * 4 * H * 2
    print(f"B={B}: plan fwd {ops.lstm_plan(B, H)} bwd {ops.lstm_plan(B, H, True)}  (slots, groups)")
    for name, (med, mn), fl in (("fwd infer", fi, 2 * flop), ("fwd train", ft, 2 * flop), ("bwd", bw, flop)):
        print(f"   {name:9s} {med * 1e3:8.1f} us median ({mn * 1e3:.1f} min)  {med * 1e3 / T:6.2f} us/step  {fl / med / 1e9:7.1f} TFLOP/s"
              f"  {B / med:8.1f} seq/ms")
    if os.environ.get("K2_TIMELINE", "1") == "1":
        print("   forward (infer) timeline, CTA 0 / slot 0 / half 0, cycles after P0 (counter seen):")
        timeline(lambda: ops.lstm_forward_fused(x, packed, B, T, False), FW, (1, 2, 3, 4, 5, 6, 7))
        print("   backward timeline, CTA 0 / slot 0 / half 0, cycles after P0 (counter seen):")
        timeline(lambda: ops.lstm_backward(packed, gates, cs, dh, B, T), BW, (2, 5, 3, 4, 7, 1, 6))
