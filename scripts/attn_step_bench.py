"""K6a alone: time per launch of the attention step kernel variants (50 launches per CUDA graph, median of 20 replays).
    python scripts/attn_step_bench.py [B T H C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rcnn_ocr_b200 as R

B, T, H, C = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (256, 64, 512, 512)
V, blank = 194, 3
L = R.lib()
g = torch.Generator(device="cuda").manual_seed(0)
projH = torch.randn(B, T, H, device="cuda", generator=g).bfloat16()
enc = torch.randn(B, T, C, device="cuda", generator=g).bfloat16()
projHf, encf = projH.float(), enc.float()
hg = torch.randn(B, H + 224, device="cuda", generator=g)
v = torch.randn(H, device="cuda", generator=g) / H ** 0.5
xcat = torch.zeros(B, C + H, dtype=torch.bfloat16, device="cuda")
probs = torch.zeros(B, 26, V, device="cuda")
y = torch.zeros(B, dtype=torch.int64, device="cuda")
lg = hg[:, H:]


def variant(name):
    s = torch.cuda.current_stream().cuda_stream
    if name == "bf16 + argmax":
        return L.rcnn_attn_step_bf16(projH.data_ptr(), hg.data_ptr(), hg.stride(0), v.data_ptr(), enc.data_ptr(), enc.stride(0),
                                     enc.stride(1), B, T, H, C, None, xcat.data_ptr(), xcat.stride(0), lg.data_ptr(), hg.stride(0),
                                     V, blank, probs[:, 3].data_ptr(), probs.stride(0), y.data_ptr(), s)
    if name == "bf16":
        return L.rcnn_attn_score_context_bf16(projH.data_ptr(), hg.data_ptr(), hg.stride(0), v.data_ptr(), enc.data_ptr(),
                                              enc.stride(0), enc.stride(1), B, T, H, C, None, xcat.data_ptr(), xcat.stride(0), s)
    if name == "fp32":
        return L.rcnn_attn_score_context_ld(projHf.data_ptr(), hg.data_ptr(), hg.stride(0), v.data_ptr(), encf.data_ptr(),
                                            encf.stride(0), encf.stride(1), B, T, H, C, None, xcat.data_ptr(), xcat.stride(0), s)
    if name == "argmax alone":
        return L.rcnn_attn_argmax_ld(lg.data_ptr(), hg.stride(0), B, V, blank, probs[:, 3].data_ptr(), probs.stride(0), y.data_ptr(), s)


for name in ("bf16 + argmax", "bf16", "fp32", "argmax alone"):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            assert variant(name) == 0
        side.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(50):
                variant(name)
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 50)
    print(f"B={B} T={T} H={H} C={C}  {name:14s} {np.median(ts):7.2f} us per launch (min {min(ts):.2f})")
