"""Determinism stress of the recurrent kernels: the same inputs through K2 forward (training variant) and K2
backward N times; hcat, the saved gates / cell states and dG must be bit-identical every time (they involve no
floating-point atomics).  Usage: python scripts/stress_block.py [B T I H] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if os.environ.get("RCNN_LIB_OVERRIDE"):      # A/B of a differently built library
    import rcnn_ocr_b200._lib as _L
    _L._LIB_PATH = os.path.abspath(os.environ["RCNN_LIB_OVERRIDE"])
from rcnn_ocr_b200 import ops

B, T, I, H = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (2500, 2, 64, 64)
iters = int(sys.argv[5]) if len(sys.argv) >= 6 else 200
g = torch.Generator(device="cuda").manual_seed(0)
k = 1.0 / H ** 0.5
ws = []
for _ in range(2):
    ws += [(torch.rand(4 * H, I, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) * k,
           (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k, (torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * k]
packed = ops.lstm_pack(*ws)
x = torch.randn(B, T, I, device="cuda", generator=g).bfloat16()
dh = torch.randn(B, T, 2 * H, device="cuda", generator=g) / (B * T) ** 0.5
from rcnn_ocr_b200 import _lib
refetch = torch.zeros(1, dtype=torch.int32, device="cuda")
_lib.lib().rcnn_debug_refetch_counter(refetch.data_ptr())
ref = None
bad = 0
for it in range(iters):
    if ops.fused_forward_supported(I, H):
        hcat, gates, cs = ops.lstm_forward_fused(x, packed, B, T, True)
    else:
        xp = ops.gemm_bf16(x.view(B * T, I), packed.wih_p, packed.bias_p, torch.float16)
        hcat, gates, cs = ops.lstm_forward(xp, packed, B, T, True)
    dG, db = ops.lstm_backward(packed, gates, cs, dh, B, T)
    cur = dict(hcat=hcat, gates=gates, csave=cs, dG=dG)
    torch.cuda.synchronize()
    if ref is None:
        ref = {k_: v.clone() for k_, v in cur.items()}
        continue
    diff = [k_ for k_, v in cur.items() if not torch.equal(v.view(torch.uint8) if v.dtype != torch.float32 else v, ref[k_].view(torch.uint8) if v.dtype != torch.float32 else ref[k_])]
    if diff:
        bad += 1
        if bad <= 8:
            msg = []
            for k_ in diff:
                a_, r_ = cur[k_].float(), ref[k_].float()
                ne = (a_ != r_) | a_.isnan()
                idx = ne.nonzero()
                first = [int(i) for i in idx[0]]
                msg.append(f"{k_}: {idx.shape[0]} elements (NaN: {int(a_.isnan().sum())}), first at {first}, "
                           f"rows b in [{int(idx[:, 0 if k_ in ('hcat', 'dG') else 2].min())}, {int(idx[:, 0 if k_ in ('hcat', 'dG') else 2].max())}]")
            print(f"iter {it}: " + "; ".join(msg))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dG, db = ops.lstm_backward(packed, gates, cs, dh, B, T)
e1.record()
torch.cuda.synchronize()
print(f"warp-level packet re-fetches after the optimistic TMA fetch: {int(refetch.item())} in {iters} forward+backward passes")
_lib.lib().rcnn_debug_refetch_counter(None)
print(f"B={B} T={T} I={I} H={H}: {bad} of {iters - 1} repeats differ; backward {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
