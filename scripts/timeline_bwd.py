import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import _lib
B, T, I, H = 256, 64, 512, 512
torch.manual_seed(0)
blk = R.BidirectionalLSTM(I, H, H).cuda()
x = torch.randn(B, T, I, device="cuda")
for it in range(3):
    y = blk(x); y.sum().backward()
tl = torch.zeros(T * 8 + T * 32, dtype=torch.int64, device="cuda")   # + per-source arrival slots (fwdx)
_lib.lib().rcnn_debug_timeline(tl.data_ptr())
y = blk(x); y.sum().backward()
torch.cuda.synchronize()
_lib.lib().rcnn_debug_timeline(None)
a = tl.cpu().numpy().reshape(T, 8).astype(np.float64)
names = ["P0 counter reached", "P1 partials stored (wait_group 0)", "M0 dG tile ready: first MMA", "M1 last MMA issued",
         "E0 accumulators complete", "E1 cell done (dG stored)", "E2 counter released", "S0 partials staged in smem"]
print("per-step intervals (cycles), median over steps 2..T-2, relative to P0 (K-local backward):")
a = a[:-1]   # the last step publishes nothing
for k in (2, 5, 3, 4, 7, 1, 6):
    d = a[2:, k] - a[2:, 0]
    print(f"  {names[k]:36s} {np.median(d):9.0f}  (min {d.min():.0f} max {d.max():.0f})")
print("  step period       ", np.median(np.diff(a[2:, 0])))
