"""Forward / backward recurrent kernels at cfg B: microseconds per launch (CUDA events) and warp-level packet
re-fetches per launch of the flag-in-data exchange.  Env: RCNN_EXCHANGE=counter|ll, RCNN_LL_DELAY=cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rcnn_ocr_b200 import ops, _lib

B, T, I, H = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (256, 64, 512, 512)
g = torch.Generator(device="cuda").manual_seed(0)
k = 1.0 / H ** 0.5
ws = []
for _ in range(2):
    ws += [(torch.rand(4 * H, I, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) * k,
           (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k]
packed = ops.lstm_pack(*ws)
x = torch.randn(B, T, I, device="cuda", generator=g).bfloat16()
dh = torch.randn(B, T, 2 * H, device="cuda", generator=g) / (B * T) ** 0.5
refetch = torch.zeros(1, dtype=torch.int32, device="cuda")
N = 30


def timed(fn, label):
    for _ in range(5):
        fn()
    refetch.zero_()
    _lib.lib().rcnn_debug_refetch_counter(refetch.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        fn()
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().rcnn_debug_refetch_counter(None)
    print(f"{label}: {e0.elapsed_time(e1) / N * 1e3:8.1f} us per launch, {refetch.item() / N:8.1f} re-fetches per launch")


tag = f"B={B} T={T} I={I} H={H} exchange={os.environ.get('RCNN_EXCHANGE', 'counter')} delay={os.environ.get('RCNN_LL_DELAY', '0')}"
timed(lambda: ops.lstm_forward_fused(x, packed, B, T, False), tag + " fwd infer")
timed(lambda: ops.lstm_forward_fused(x, packed, B, T, True), tag + " fwd train")
hcat, gates, cs = ops.lstm_forward_fused(x, packed, B, T, True)
timed(lambda: ops.lstm_backward(packed, gates, cs, dh, B, T), tag + " bwd      ")
