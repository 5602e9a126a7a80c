"""One small launch of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck`.
Shapes are small (the tools slow a kernel down 10-1000x) but keep several work items per CTA group, so the
recurrent kernels' step exchange, the item loop and the weight re-staging are all exercised:

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

The logs are kept under profiles/ (sanitizer_*_rNN.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import ops

which = set(sys.argv[1:]) or {"lstm", "gemm", "ctc", "decode", "metrics", "attn"}
torch.manual_seed(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(0)

if "lstm" in which:
    # block: fused forward (training mode), backward, weight gradients; B > 64 * groups so that groups loop over items
    for (B, T, I, H) in ((96, 4, 64, 128), (70, 3, 128, 64)):
        blk = R.BidirectionalLSTM(I, H, H).to(dev)
        x = torch.randn(B, T, I, device=dev, generator=g).requires_grad_(True)
        y = blk(x)
        y.square().mean().backward()
        blk.eval()
        with torch.no_grad():
            y2 = blk(x)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all() and torch.isfinite(x.grad).all() and torch.isfinite(y2).all()
        print(f"lstm B={B} T={T} I={I} H={H}: ok", flush=True)

if "gemm" in which:
    for (M, N, K) in ((300, 195, 128), (512, 256, 192), (1024, 512, 256)):
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        b = torch.randn(N, K, device=dev, generator=g).bfloat16()
        bias = torch.randn(N, device=dev, generator=g)
        o = ops.gemm_bf16(a, b, bias)
        o2 = ops.gemm_bf16_atb(a, torch.randn(M, N, device=dev, generator=g).bfloat16())
        torch.cuda.synchronize()
        assert torch.isfinite(o).all() and torch.isfinite(o2).all()
        print(f"gemm M={M} N={N} K={K}: ok", flush=True)

if "ctc" in which:
    T, N, C = 16, 24, 195
    x = torch.randn(N, T, C, device=dev, generator=g).requires_grad_(True)
    tl = torch.randint(0, 7, (N,), device=dev, generator=g)
    tg = torch.randint(1, C, (N, 6), device=dev, generator=g)
    il = torch.randint(8, T + 1, (N,), device=dev, generator=g)
    loss = R.ctc_loss_from_logits(x.permute(1, 0, 2), tg, il, tl, 0, "mean", True)
    loss.backward()
    lp = torch.log_softmax(x.detach(), 2).permute(1, 0, 2).requires_grad_(True)
    l2 = R.CTCLoss(0, "sum", False)(lp, tg, il, tl)
    l2.backward()
    torch.cuda.synchronize()
    print(f"ctc: loss {loss.item():.4f} ok", flush=True)

if "decode" in which:
    for dt in (torch.float32, torch.bfloat16):
        lg = torch.randn(37, 16, 195, device=dev, generator=g).to(dt)
        ids, lens = R.ctc_greedy_ids(lg)
        torch.cuda.synchronize()
    print("decode: ok", flush=True)

if "metrics" in which:
    lg = torch.randn(9, 16, 195, device=dev, generator=g)
    ids, lens = R.ctc_greedy_ids(lg)
    tl = torch.randint(1, 7, (9,), device=dev, generator=g)
    tg = torch.randint(1, 195, (9, 6), device=dev, generator=g)
    itos = [chr(0x430 + i % 32) if i % 7 else " " for i in range(194)]
    table = R.CharsetTable(itos, dev)
    for words in (False, True):
        d, nr, nh = R.edit_stats(ids, lens, tg, tl, table, words=words)
    torch.cuda.synchronize()
    print("metrics: ok", flush=True)

if "attn" in which:
    from rcnn_ocr_b200 import attention
    att = attention.Attention(64, 64, 30, sos_id=1, eos_id=2, pad_id=0, blank_id=3).to(dev).eval()
    bh = torch.randn(5, 8, 64, device=dev, generator=g)
    with torch.no_grad():
        out = att(bh, None, is_train=False, batch_max_length=4)
    torch.cuda.synchronize()
    print("attn: ok", flush=True)
print("sanitize_small: done", flush=True)
