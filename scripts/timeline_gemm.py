"""Per-tile clock64 timeline of CTA 0 of the CTA-pair GEMM on the xp-projection shape (M=16384, N=4096, K=512)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from rcnn_ocr_b200 import _lib, ops
M, N, K = 16384, 4096, 512
if len(sys.argv) > 3:
    M, N, K = [int(v) for v in sys.argv[1:4]]
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
for _ in range(3):
    ops.gemm_bf16(a, b, bias, torch.float16)
tl = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
_lib.lib().rcnn_debug_timeline(tl.data_ptr())
ops.gemm_bf16(a, b, bias, torch.float16)
torch.cuda.synchronize()
_lib.lib().rcnn_debug_timeline(None)
t = tl.cpu().numpy().reshape(64, 8).astype(np.float64)
t = t[(t[:, 0] > 0) & (t[:, 3] > 0)]
names = ["M0 tile start", "M1 accumulator free", "M2 first stage full", "M3 last MMA + commit issued", "E4 accumulator complete (epilogue)", "E5 drained"]
print(f"{M}x{N}x{K}: {len(t)} tiles by CTA 0; cycles relative to the tile start (median over tiles 2..)")
names += ["E6 first chunk read from TMEM", "E7 first chunk handed to TMA"]
for k in (1, 2, 3, 4, 6, 7, 5):
    d = t[2:, k] - t[2:, 0]
    print(f"  {names[k]:36s} {np.median(d):9.0f}  (min {d.min():.0f} max {d.max():.0f})")
print("  tile period", np.median(np.diff(t[2:, 0])))
