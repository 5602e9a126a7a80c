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
tl = torch.zeros(T * 8 + T * 32 + 256, dtype=torch.int64, device="cuda")
_lib.lib().rcnn_debug_timeline(tl.data_ptr())
y = blk(x); y.sum().backward()
torch.cuda.synchronize()
_lib.lib().rcnn_debug_timeline(None)
a = tl.cpu().numpy()
p0 = a[5 * 8 + 0]
ch = a[T * 8: T * 8 + 32] - p0
print("chunk arrival (cycles after P0) at step 5:", ch.tolist())
print("deltas:", np.diff(ch).tolist())
bw = a[T * 8 + 64: T * 8 + 96] - p0
print("wait durations (after-before):", (ch - bw).tolist())
af = a[T * 8 + 160: T * 8 + 192] - p0
am = a[T * 8 + 192: T * 8 + 224] - p0
ac = a[T * 8 + 224: T * 8 + 256] - p0
print("fence:", (af - ch).tolist()[:12])
print("4 MMAs:", (am - af).tolist()[:12])
print("commit:", (ac - am).tolist()[:12])
print("loop back to next wait:", (bw[1:] - ac[:-1]).tolist()[:12])
pi = a[T * 8 + 128: T * 8 + 160] - p0
print("producer issue times:", pi.tolist())
print("issue->arrival latency:", (ch - pi).tolist())
