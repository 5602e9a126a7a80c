"""Per-timestep clock64 timeline of CTA 0 / half 0 of the fused forward kernel (lstm_fwdx_kernel), cfg B."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import _lib
B, T, I, H = 256, 64, 512, 512
torch.manual_seed(0)
blk = R.BidirectionalLSTM(I, H, H).cuda()
x = torch.randn(B, T, I, device="cuda")
LL = os.environ.get("RCNN_EXCHANGE", "counter") == "ll"
names = {1: "M0 first h box landed", 2: "M1 half-0 MMAs issued", 3: "M2 half-1 MMAs issued", 4: "E0 half-0 accumulator complete",
         5: "E1 half-0 cell phase done, h stored",
         6: "R0 first TMA box of the NEXT tile landed" if LL else "R0 publisher saw h_staged",
         7: "R1 ... and validated by warp 2" if LL else "R1 release issued (MEMBAR + RED done)"}
print("exchange:", "flag-in-data (P0 = half 0's canary words seen by the loader warp)" if LL else "release/acquire counter (P0 = counter seen)")
for mode in ("train", "infer"):
    def run():
        if mode == "train":
            return blk(x.requires_grad_(True))
        with torch.no_grad():
            return blk(x)
    for it in range(3):
        run()
    tl = torch.zeros(T * 8 + T * 32, dtype=torch.int64, device="cuda")
    _lib.lib().rcnn_debug_timeline(tl.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.lib().rcnn_debug_timeline(None)
    raw = tl.cpu().numpy().astype(np.float64)
    a = raw[:T * 8].reshape(T, 8)
    src = raw[T * 8:].reshape(T, 32)
    print(mode, "forward (fused): per-step cycles, median over steps 2..T-2, relative to P0 (half 0's counter seen):")
    for k in range(1, 8):
        d = a[2:T - 1, k] - a[2:T - 1, 0]
        print(f"  {names[k]:40s} {np.median(d):9.0f}  (min {d.min():.0f} max {d.max():.0f})")
    print("  step period       ", np.median(np.diff(a[2:T - 1, 0])))
    if LL:
        # canary of loader iteration s = h stored in step s-1: arrival relative to CTA 0's own E1 of step s-1
        d = src[3:T - 1, :] - a[2:T - 2, 5:6]
        print("  canary arrival of source CTA j (half 0: lanes 0-15, half 1: 16-31) minus CTA 0's own half-0 h store, median cycles:")
        print("   ", " ".join(f"{np.median(d[:, j]):.0f}" for j in range(32)))
