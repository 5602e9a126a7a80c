"""Size sweeps of the HBM-bound kernels (SURVEY.md section 8d cfg 2 and cfg 5): fused CTC loss fwd+bwd over
N, greedy decode over B (fp32 and bf16 logits).  CUDA-event timing, inputs rotate over buffers
that together exceed L2.  Prints one JSON object; committed under profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcnn_ocr_b200 as R
from rcnn_ocr_b200 import _lib
import bench

T, C = 64, 195
peaks = bench.load_peaks()
dev = torch.device("cuda", 0)


def time_ms(fn, nbuf, reps=30, warm=5):
    for i in range(warm):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_ms(fn, nbuf, kid, reps=20):
    _lib.prof_enable(True)
    for i in range(reps):
        fn(i % nbuf)
    torch.cuda.synchronize()
    ms, n = _lib.prof_read(kid)
    _lib.prof_enable(False)
    return ms / max(n, 1)


out = {"peaks": peaks, "ctc": [], "decode": []}
g = torch.Generator(device="cuda").manual_seed(0)
for N in (256, 1024, 4096, 16384):
    per = T * N * C * 4
    nbuf = max(2, min(32, (300 << 20) // (2 * per) + 1))
    xs = [torch.randn(N, T, C, device=dev, generator=g).requires_grad_(True) for _ in range(nbuf)]
    tl = torch.randint(1, 33, (N,), device=dev, generator=g)
    tg = torch.randint(1, C, (N, 32), device=dev, generator=g)
    il = torch.full((N,), T, device=dev)

    def step(i):
        x = xs[i]
        x.grad = None
        loss = R.ctc_loss_from_logits(x.permute(1, 0, 2), tg, il, tl, 0, "mean", True, max_target_length=32)
        loss.backward()

    ms_api = time_ms(step, nbuf)
    ms_k = kernel_ms(step, nbuf, 1)
    alg = 2.0 * T * C * 4 * N
    out["ctc"].append({"N": N, "ms_api_fwd_bwd": round(ms_api, 4), "ms_kernel": round(ms_k, 4),
                       "seq_per_s_kernel": round(N / (ms_k * 1e-3)), "GBps_kernel": round(alg / ms_k / 1e6, 1),
                       "hbm_frac_kernel": round(alg / ms_k / 1e6 / peaks["hbm"], 4), "buffers": nbuf})
    del xs

for dtype, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    for B in (1, 8, 64, 256, 1024, 4096, 16384):
        per = T * B * C * (4 if dtype == torch.float32 else 2)
        nbuf = max(2, min(64, (300 << 20) // per + 1))
        xs = [torch.randn(B, T, C, device=dev, generator=g).to(dtype) for _ in range(nbuf)]

        def step(i):
            R.ctc_greedy_ids(xs[i])

        ms_api = time_ms(step, nbuf)
        ms_k = kernel_ms(step, nbuf, 0)
        alg = float(per + B * (T + 1) * 4)
        out["decode"].append({"dtype": name, "B": B, "ms_api": round(ms_api, 4), "ms_kernel": round(ms_k, 4),
                              "seq_per_s_kernel": round(B / (ms_k * 1e-3)), "GBps_kernel": round(alg / ms_k / 1e6, 1),
                              "hbm_frac_kernel": round(alg / ms_k / 1e6 / peaks["hbm"], 4), "buffers": nbuf})
        del xs
# ---- BASELINE configs[4]: greedy inference throughput over the batch size (encoder 2xBiLSTM(512) + head + decode,
# eval mode, device-resident features), eager launches and one CUDA-graph replay per batch; plus the train step.
out["inference"], out["train"] = [], []
step = bench.TrainStep(dev, 1, 0)
infer = bench.InferStep(step)
step.enc.eval(); step.head.eval()
for B in (1, 8, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    nbuf = max(2, min(8, (160 << 20) // (B * T * 512 * 4) + 1))
    feats = [torch.randn(B, T, 512, device=dev) for _ in range(nbuf)]
    ms_eager = time_ms(lambda i: infer.device(feats[i]), nbuf, reps=20)
    gi = R.GraphedStep(infer.device, [feats[0]])
    ms_graph = time_ms(lambda i: gi(feats[i]), nbuf, reps=20)
    out["inference"].append({"B": B, "ms_eager": round(ms_eager, 4), "ms_graph": round(ms_graph, 4),
                             "lines_per_s_graph": round(B / (ms_graph * 1e-3), 1)})
    del gi, feats
step.enc.train(); step.head.train()
_warm = [t.to(dev) for t in bench.make_batch(256, 5)]
for _ in range(5):            # one-time costs (library workspaces, attribute calls, allocator growth) stay out of the first row
    step(*_warm)
torch.cuda.synchronize()
for B in (64, 128, 256, 512, 1024):
    batches = [[t.to(dev) for t in bench.make_batch(B, 77 + i)] for i in range(3)]
    ms = time_ms(lambda i: step(*batches[i]), 3, reps=20, warm=5)
    out["train"].append({"B": B, "ms_eager": round(ms, 4), "lines_per_s": round(B / (ms * 1e-3), 1)})
print(json.dumps(out, indent=1))
