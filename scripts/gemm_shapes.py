"""CUDA-event time of every tn-GEMM shape of the cfg-B train step (one launch each, inputs L2-cold by rotation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rcnn_ocr_b200 import ops
shapes = [("xp (f16 out)", 16384, 4096, 512, torch.float16, True), ("linear blk1 (bf16 out)", 16384, 512, 1024, torch.bfloat16, True),
          ("linear blk2 (f32 out)", 16384, 512, 1024, torch.float32, True), ("head (f32, N=195)", 16384, 195, 512, torch.float32, True),
          ("d_enc = dlogits W_head", 16384, 512, 200, torch.float32, False), ("dhcat (f32 out)", 16384, 1024, 512, torch.float32, False),
          ("dX (bf16 out)", 16384, 512, 4096, torch.bfloat16, False)]
for name, M, N, K, dt, has_bias in shapes:
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(4)]
    Bm = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda") if has_bias else None
    outs = [torch.empty(M, N, device="cuda", dtype=dt) for _ in range(4)]
    for i in range(4):
        ops.gemm_bf16(As[i], Bm, bias, dt, out=outs[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ops.gemm_bf16(As[i % 4], Bm, bias, dt, out=outs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name:28s} M={M} N={N} K={K}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
