#!/usr/bin/env python
"""Benchmark of the RCNN-OCR sequence-recognition hot path on B200 (BASELINE.json metric:
text-lines/sec for the train step and greedy inference; % of roofline).

    python bench.py --gpus N --steps K --warmup W            # ours (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path (port)

Workload ("cfgB", BASELINE.json configs[1]+[2]): per GPU 256 synthetic text lines,
T=64 feature columns of 512 channels -> 2 x BidirectionalLSTM(hidden 512) -> CTC head (195
classes = configs/charset.txt + blank) -> fused log_softmax + CTC loss (labels U{1..32}) ->
backward -> NCCL gradient all-reduce (N > 1) -> Adam step.  A "step" is one such pass over one
batch; weak scaling (per-GPU batch fixed).  The SE-ResNet31 backbone is outside the hot path
 and is not part of the timed region: the inputs are its feature columns.

One JSON line on rank 0; see DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = dict(T=64, IN=512, H=512, C=195, LMAX=32, B=256)
L2_BYTES = 126 * 1024 * 1024


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (None if absent)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kernel)
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------- synthetic data
def make_batch(B, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, CFG["T"], CFG["IN"], generator=g)
    tl = torch.randint(1, CFG["LMAX"] + 1, (B,), generator=g)
    tg = torch.randint(1, CFG["C"], (B, CFG["LMAX"]), generator=g)
    il = torch.full((B,), CFG["T"], dtype=torch.long)
    out = [feats, tg, il, tl]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- CPU baseline (port)
def cpu_reference(steps, warmup, sample_B, workload="train"):
    """The reference's own torch-CPU op sequence for the path (oracle/ref_port.py), all host
    threads, on a bounded sample of the workload.  Returns (lines/s, ms/step, cores)."""
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc = ref_port.make_encoder(CFG["IN"], CFG["H"])
    head = torch.nn.Linear(CFG["H"], CFG["C"])
    feats, tg, il, tl = make_batch(sample_B, 1234)
    alphabet = [chr(0x4E00 + i) for i in range(CFG["C"] - 1)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if workload == "train":
            ref_port.train_step(enc, head, feats, tg, il, tl)
        else:
            ref_port.infer_step(enc, head, feats, alphabet)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return sample_B / (ms / 1e3), ms, cores


def run_reference(args, rank):
    if rank != 0:
        return
    sample_B = args.ref_batch
    val, ms, cores = cpu_reference(args.steps, max(args.warmup, 1), sample_B)
    line = {
        "impl": "reference", "metric": "text-lines/sec", "value": round(val, 2), "unit": "lines/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_B, 1),
        "cpu_baseline": {"value": round(val, 2), "unit": "lines/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_B} lines/step x {args.steps} steps of the same T/H/C workload "
                                   "(oracle/ref_port.py: the reference's torch-CPU op sequence)"},
        "e2e": {"value": round(val, 2), "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, B, world):
    return {"workload": "cfgB train step: 2xBiLSTM(512)+CTC head+fused CTC loss fwd/bwd+Adam "
                        "(BASELINE configs[1]+[2]); infer: same encoder + greedy CTC decode",
            "per_gpu_batch": B, "global_batch": B * world, "T": CFG["T"], "in": CFG["IN"], "hidden": CFG["H"],
            "classes": CFG["C"], "label_len": "U{1..32}", "parallelism": f"dp{world}", "feature_dtype": "bf16",
            "l2_policy": "inputs rotate over a ring larger than L2 (8 sets x 33.5 MB features + 8 gradient/"
                         "activation sets)"}


# ----------------------------------------------------------------------------- ours
class TrainStep:
    def __init__(self, device, world, rank):
        import rcnn_ocr_b200 as R
        from rcnn_ocr_b200.dist import GradAllReducer
        self.R = R
        torch.manual_seed(0)
        self.enc = R.make_enc_rnn(CFG["IN"], CFG["H"], out_dtype=torch.bfloat16).to(device)   # feeds the bf16 head GEMM
        self.head = R.CTCHead(CFG["H"], CFG["C"]).to(device)
        self.params = list(self.enc.parameters()) + list(self.head.parameters())
        self.opt = torch.optim.Adam(self.params, lr=5.1e-4, weight_decay=1.95e-5, fused=True, capturable=True)
        self.reducer = GradAllReducer(self.params, producers=[list(m.parameters()) for m in self.enc.children()]
                                      + [list(self.head.parameters())]) if world > 1 else None
        self.world = world

    def __call__(self, feats, tg, il, tl):
        self.opt.zero_grad(set_to_none=True)
        # the feature columns come from a trainable backbone (model/model.py:216-219): they need their gradient,
        # so block 1's dX GEMM is part of the step (backward = 2 x forward FLOPs)
        feats = feats.detach().requires_grad_(True)
        logits = self.head(self.enc(feats))                                  # [B,T,C] fp32
        loss = self.R.ctc_loss_from_logits(logits.permute(1, 0, 2), tg, il, tl, 0, "mean", True,
                                           max_target_length=CFG["LMAX"])
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss


class InferStep:
    def __init__(self, train: TrainStep):
        self.t = train
        self.alphabet = [chr(0x4E00 + i) for i in range(CFG["C"] - 1)]

    @torch.no_grad()
    def device(self, feats):
        logits = self.t.head(self.t.enc(feats))
        return self.t.R.ctc_greedy_ids(logits)

    @torch.no_grad()
    def e2e(self, feats):
        feats = feats.to(self.t.params[0].device, non_blocking=True)   # pinned host -> device
        logits = self.t.head(self.t.enc(feats))
        return self.t.R.ctc_greedy_decoder(logits, self.alphabet, batch_first=True)



# ----------------------------------------------------------------------------- vendor bar (torch on the same B200)
class _VendorBlock(torch.nn.Module):
    """model/model.py:151-163 as the reference writes it: nn.LSTM (cuDNN on CUDA) + nn.Linear (cuBLAS)."""

    def __init__(self, n_in, n_hidden, n_out):
        super().__init__()
        self.rnn = torch.nn.LSTM(n_in, n_hidden, bidirectional=True, batch_first=True)
        self.linear = torch.nn.Linear(2 * n_hidden, n_out)

    def forward(self, x):
        self.rnn.flatten_parameters()
        y, _ = self.rnn(x)
        return self.linear(y)


class VendorStep:
    """The same cfg-B train step through torch's own CUDA kernels (cuDNN LSTM, cuBLAS, ATen log_softmax + CTC,
    fused Adam) under bf16 autocast, as training/train.py:499-508 runs its model -- the bar to beat on this box."""

    def __init__(self, device, dtype=torch.bfloat16):
        torch.manual_seed(0)
        self.dtype = dtype
        self.enc = torch.nn.Sequential(_VendorBlock(CFG["IN"], CFG["H"], CFG["H"]),
                                       _VendorBlock(CFG["H"], CFG["H"], CFG["H"])).to(device)
        self.head = torch.nn.Linear(CFG["H"], CFG["C"]).to(device)
        self.params = list(self.enc.parameters()) + list(self.head.parameters())
        self.opt = torch.optim.Adam(self.params, lr=5.1e-4, weight_decay=1.95e-5, fused=True, capturable=True)
        self.il = [CFG["T"]] * CFG["B"]

    def __call__(self, feats, tg, il, tl):
        import torch.nn.functional as F
        self.opt.zero_grad(set_to_none=True)
        feats = feats.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=self.dtype):
            logits = self.head(self.enc(feats))
        lp = F.log_softmax(logits.float(), dim=2).permute(1, 0, 2)
        loss = F.ctc_loss(lp, tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
        loss.backward()
        self.opt.step()
        return loss

    def no_ctc(self, feats, w):
        """The same step with the CTC loss replaced by a weighted sum of the logits: everything but ATen's CTC, which
        cannot be captured into a CUDA graph (it stages its offsets through pageable host memory)."""
        self.opt.zero_grad(set_to_none=True)
        feats = feats.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=self.dtype):
            logits = self.head(self.enc(feats))
        loss = (logits.float() * w).sum()
        loss.backward()
        self.opt.step()
        return loss

    def ctc_only(self, logits, tg, il, tl):
        import torch.nn.functional as F
        logits.grad = None
        lp = F.log_softmax(logits, dim=2).permute(1, 0, 2)
        loss = F.ctc_loss(lp, tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
        loss.backward()
        return loss

    @torch.no_grad()
    def infer(self, feats):
        with torch.autocast("cuda", dtype=self.dtype):
            logits = self.head(self.enc(feats))
        return logits.argmax(dim=2)     # (the reference then loops over B*T .item() calls on the host; not timed here)


def run_vendor(args, dev, B, timed_fn):
    """Returns the `vendor` object: torch-CUDA train step and encoder+argmax inference at cfg B, device-resident,
    CUDA-graph replay when torch's ops allow the capture (else eager, stated).  Measured under bf16 autocast (this
    build's arithmetic) and under fp16 autocast (what the reference's torch.cuda.amp.autocast() picks,
    training/train.py:499); `value` is the faster of the two -- the bar to beat."""
    import rcnn_ocr_b200 as R
    # F.ctc_loss takes the lengths from the host (IntArrayRef): CPU tensors avoid a sync per step
    # (a captured graph bakes them in, so the ring rotates the features and keeps batch 0's targets / lengths)
    il0, tl0 = dev[0][2].cpu(), dev[0][3].cpu()
    batches = [[b[0], dev[0][1], il0, tl0] for b in dev]
    runs = {}
    for name, dtype in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        v = VendorStep(dev[0][0].device, dtype)
        mode = "eager"
        fn = lambda i: v(*batches[i % len(batches)])
        if args.graph:
            try:
                g = R.GraphedStep(lambda f, t: v(f, t, il0, tl0), batches[0][:2])
                fn = lambda i: g(*batches[i % len(batches)][:2])
                mode = "cuda-graph replay"
            except Exception as e:  # noqa: BLE001 -- ATen's CTC stages its offsets through pageable host memory
                torch.cuda.synchronize()
                mode = f"eager (capture failed: {type(e).__name__})"
        ms = timed_fn(fn)
        # launch-overhead-free estimate of the same step: (everything but the CTC) as one graph replay + the CTC part alone
        est = None
        if args.graph and mode != "cuda-graph replay":
            try:
                wsum = torch.randn(B, CFG["T"], CFG["C"], device=dev[0][0].device) / (B * CFG["T"])
                g2 = R.GraphedStep(lambda f: v.no_ctc(f, wsum), [batches[0][0]])
                ms_g = timed_fn(lambda i: g2(batches[i % len(batches)][0]))
                lg = torch.randn(B, CFG["T"], CFG["C"], device=dev[0][0].device, requires_grad=True)
                ms_c = timed_fn(lambda i: v.ctc_only(lg, dev[0][1], il0, tl0))
                est = {"ms_per_step": round(ms_g + ms_c, 4), "graph_without_ctc_ms": round(ms_g, 4), "ctc_alone_eager_ms": round(ms_c, 4),
                       "value": round(B / ((ms_g + ms_c) * 1e-3), 1)}
                del g2
            except Exception as e:  # noqa: BLE001
                torch.cuda.synchronize()
                est = {"error": type(e).__name__}
        v.enc.eval()
        ginf, mode_inf = v.infer, "eager"
        if args.graph:
            try:
                ginf = R.GraphedStep(v.infer, [batches[0][0]])
                mode_inf = "cuda-graph replay"
            except Exception:  # noqa: BLE001
                torch.cuda.synchronize()
        ms_inf = timed_fn(lambda i: ginf(batches[i % len(batches)][0]))
        if est and "value" in est and est["value"] > B / (ms * 1e-3):
            ms, mode = est["ms_per_step"], "cuda-graph replay of everything but the CTC + ATen CTC eager (sum of the two)"
        runs[name] = {"value": round(B / (ms * 1e-3), 1), "ms_per_step": round(ms, 4), "launch_mode": mode, "split": est,
                      "infer": {"value": round(B / (ms_inf * 1e-3), 1), "ms_per_step": round(ms_inf, 4), "launch_mode": mode_inf}}
        del v, ginf
    best = max(runs, key=lambda k: runs[k]["value"])
    best_inf = max(runs, key=lambda k: runs[k]["infer"]["value"])
    return {"value": runs[best]["value"], "unit": "lines/s", "ms_per_step": runs[best]["ms_per_step"],
            "launch_mode": runs[best]["launch_mode"], "autocast": best,
            "infer": {"value": runs[best_inf]["infer"]["value"], "ms_per_step": runs[best_inf]["infer"]["ms_per_step"],
                      "autocast": best_inf, "launch_mode": runs[best_inf]["infer"]["launch_mode"],
                      "note": "encoder + head + argmax only (no collapse / strings)"},
            "by_autocast_dtype": runs,
            "what": "torch 2.11 CUDA: nn.LSTM (cuDNN, bidirectional, batch_first) + nn.Linear x2 + head under autocast, "
                    "F.log_softmax + F.ctc_loss (ATen CUDA), backward, fused Adam -- model/model.py:154-162, "
                    "training/train.py:499-508 on the same GPU, same inputs, device-resident"}


# ----------------------------------------------------------------------------- cfg 1: minimal_inference
def run_cfg1(args, device, timed_fn, cpu: bool):
    """BASELINE configs[0] (minimal_inference.py:13-15 -> inference.py:155-180): RCNN(194, hidden 256), 32 synthetic
    lines of 32x128, eval, greedy decode -> strings."""
    import rcnn_ocr_b200 as R
    torch.manual_seed(0)
    model = R.RCNN(194, hidden_size=256).to(device).eval().to(memory_format=torch.channels_last)
    alphabet = [chr(0x4E00 + i) for i in range(194)]
    g = torch.Generator().manual_seed(1234)
    host = [(torch.rand(32, 3, 32, 128, generator=g) * 2 - 1).pin_memory() for _ in range(4)]
    dev = [h.to(device).contiguous(memory_format=torch.channels_last) for h in host]

    if os.environ.get("RCNN_FOLD_BACKBONE", "1") == "1":
        model.fold_backbone(torch.bfloat16)                 # BatchNorm folded, conv + bias + ReLU fused (model.FoldedBackbone)

    @torch.no_grad()
    def fwd(x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feats = model._features(x)
        return R.ctc_greedy_ids(model.ctc_head(model._encode_features(feats)))

    gfwd, mode = fwd, "eager"
    if args.graph:
        try:
            gfwd = R.GraphedStep(fwd, [dev[0]])
            mode = "cuda-graph replay"
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            mode = f"eager (capture failed: {type(e).__name__})"
    ms = timed_fn(lambda i: gfwd(dev[i % 4]))

    def e2e(i):
        x = host[i % 4].to(device, non_blocking=True).contiguous(memory_format=torch.channels_last)
        ids, lens = gfwd(x)
        return R.ids_to_text(ids, lens, alphabet)[0]

    ms_e2e = timed_fn(e2e)
    # the same model the way the reference runs it on a GPU (inference.py:155-166): fp32, NCHW, eager launches -- what the
    # channels_last + bf16 autocast + graph replay of the backbone (section 8f-2) is measured against
    model_ref = R.RCNN(194, hidden_size=256).to(device).eval()
    xs_ref = [h.to(device) for h in host]

    @torch.no_grad()
    def fwd_ref(i):
        return R.ctc_greedy_ids(model_ref(xs_ref[i % 4], is_train=False))

    ms_fp32 = timed_fn(fwd_ref)
    del model_ref, xs_ref
    # from decoded pixels: uint8 line images of assorted sizes -> K7 (resize + pad + normalise + batch, one launch) ->
    # model -> strings, against the reference's host-side input step (cv2 + numpy per image, inference.py:93-124,159-164)
    rng = np.random.default_rng(7)
    raw = [[rng.integers(0, 256, (int(rng.integers(24, 80)), int(rng.integers(90, 400)), 3), dtype=np.uint8) for _ in range(32)]
           for _ in range(4)]
    pre = R.LinePreprocessor(32, 128, device)

    def e2e_pixels(i):
        x = pre(raw[i % 4]).contiguous(memory_format=torch.channels_last)
        ids, lens = gfwd(x)
        return R.ids_to_text(ids, lens, alphabet)[0]

    ms_pix = timed_fn(e2e_pixels)
    ms_k7 = timed_fn(lambda i: pre(raw[i % 4]))
    host_prep_ms = None
    if cpu:
        from oracle import host_ref
        t0 = time.perf_counter()
        for k in range(3):
            torch.stack([torch.from_numpy(host_ref.normalize_chw(host_ref.resize_and_pad(im, 32, 128))) for im in raw[k]]).to(device)
        torch.cuda.synchronize()
        host_prep_ms = (time.perf_counter() - t0) / 3 * 1e3
    T = int(model._features(dev[0]).shape[1])
    # the encoder's TRAIN step at this shape (T = 16, H = 256, 32 lines; features given): T does not fill the 64-step K chunks
    # of rcnn_lstm_weight_grads, so the weight gradients take the h_prev-copy + grouped-product + unpack route (ops.py)
    from rcnn_ocr_b200 import ops as _ops
    enc_a = R.make_enc_rnn(512, 256).to(device)
    head_a = R.CTCHead(256, 195).to(device)
    opt_a = torch.optim.Adam(list(enc_a.parameters()) + list(head_a.parameters()), lr=5.1e-4, fused=True, capturable=True)
    ga = torch.Generator().manual_seed(77)
    fa = [torch.randn(32, T, 512, generator=ga).to(device) for _ in range(4)]
    tga = torch.randint(1, 195, (32, 8), generator=ga).to(device)
    ila, tla = torch.full((32,), T, dtype=torch.int64, device=device), torch.randint(1, 7, (32,), generator=ga).to(device)

    def train_a(f):
        opt_a.zero_grad(set_to_none=True)
        loss = R.ctc_loss_from_logits(head_a(enc_a(f.detach().requires_grad_(True))).permute(1, 0, 2), tga, ila, tla, 0, "mean", True,
                                      max_target_length=8)
        loss.backward()
        opt_a.step()
        return loss

    g_a = R.GraphedStep(train_a, [fa[0]]) if args.graph else train_a
    ms_ta = timed_fn(lambda i: g_a(fa[i % 4]))
    train_shape = {"value": round(32 / (ms_ta * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms_ta, 4),
                   "weight_grads_fast_path": bool(_ops.weight_grads_supported(512, 256, T)),
                   "note": "encoder (2 x BiLSTM 256) + CTC head + fused CTC + backward + Adam on [32, %d, 512] features, graph replay" % T}
    del enc_a, head_a, opt_a, g_a
    # the reference's LIVE model at this configuration: the same backbone and encoder with its attention decoder
    # (RCNN(decoder="attention"): model/model.py:223-227 -> Attention._greedy_decode, 26 steps) -> token ids
    model_at = R.RCNN(194, hidden_size=256, decoder="attention").to(device).eval().to(memory_format=torch.channels_last)
    if os.environ.get("RCNN_FOLD_BACKBONE", "1") == "1":
        model_at.fold_backbone(torch.bfloat16)

    @torch.no_grad()
    def fwd_at(x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feats = model_at._features(x)
        enc = model_at._encode_features(feats)
        return model_at.attn(enc.float(), is_train=False, batch_max_length=25).argmax(dim=-1)

    g_at = fwd_at
    if args.graph:
        try:
            g_at = R.GraphedStep(fwd_at, [dev[0]])
        except Exception:  # noqa: BLE001
            torch.cuda.synchronize()
    ms_at = timed_fn(lambda i: g_at(dev[i % 4]))
    attn_shape = {"value": round(32 / (ms_at * 1e-3), 1), "unit": "lines/s", "ms_per_batch": round(ms_at, 4),
                  "note": "RCNN(194, hidden 256, decoder='attention') eval: backbone -> encoder -> the reference's attention decoder "
                          "(greedy, 26 steps) -> token ids, device-resident, " + ("graph replay" if g_at is not fwd_at else "eager")}
    del model_at, g_at
    out = {"value": round(32 / (ms * 1e-3), 1), "unit": "lines/s", "ms_per_batch": round(ms, 4), "launch_mode": mode,
           "encoder_train_step": train_shape, "attention_decoder": attn_shape,
           "e2e": {"value": round(32 / (ms_e2e * 1e-3), 1), "ms_per_batch": round(ms_e2e, 4),
                   "h2d_bytes_per_step": host[0].numel() * 4, "d2h_bytes_per_step": 32 * (T + 1) * 4},
           "e2e_from_pixels": {"value": round(32 / (ms_pix * 1e-3), 1), "ms_per_batch": round(ms_pix, 4),
                               "input_step_ms": round(ms_k7, 4), "host_input_step_ms": None if host_prep_ms is None else round(host_prep_ms, 3),
                               "h2d_bytes_per_step": int(sum(im.size for im in raw[0])),
                               "note": "32 uint8 line images (24-80 x 90-400 px) -> K7 on the device (one pinned copy + one launch) -> "
                                       "model -> strings; host_input_step_ms = the reference's per-image cv2 resize + pad + "
                                       "normalise + stack + copy on this host (oracle/host_ref, wall clock)"},
           "backbone_fp32_nchw_eager": {"value": round(32 / (ms_fp32 * 1e-3), 1), "ms_per_batch": round(ms_fp32, 4),
                                        "note": "same RCNN, backbone in fp32 / NCHW / eager launches (the reference's GPU settings), "
                                                "device-resident; `value` above uses channels_last + bf16 autocast + graph replay"},
           "config": {"workload": "cfg1 minimal_inference: RCNN(194, hidden 256) eval, x[32,3,32,128] in [-1,1], greedy "
                                  "CTC decode -> strings", "T": T, "backbone": "torch/cuDNN, channels_last + bf16, BatchNorm folded into the convolutions, conv + bias + ReLU "
                                  "as one cuDNN call (model.FoldedBackbone), graph replay"}}
    if cpu:
        from oracle import ref_port
        torch.set_num_threads(os.cpu_count() or 1)
        ref = ref_port.RefRCNN(194, 256).eval()
        x = host[0].clone()
        ref.infer(x[:4], alphabet)
        t0 = time.perf_counter()
        ref.infer(x, alphabet)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": round(32 / dt, 1), "unit": "lines/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "one batch of 32 lines after a 4-line warm-up, oracle/ref_port.RefRCNN (SE-ResNet31 + "
                                         "nn.LSTM encoder + head on torch CPU, the reference's per-frame .item() decode loop)"}
    return out


# ----------------------------------------------------------------------------- cfg 4: full train step, strong scaling
class FullTrainStep:
    """BASELINE configs[3] (training/train.py:493-508 with configs/config.json:20-29): images -> SE-ResNet31 (torch/cuDNN,
    channels_last, bf16 autocast) -> enc_rnn (K1/K2) -> CTC head -> fused CTC loss (K3) -> backward -> gradient
    all-reduce of ALL parameters -> Adam.  GLOBAL batch 512, sharded over the ranks (strong scaling)."""
    GLOBAL_B = 512

    def __init__(self, device, world, hidden):
        import rcnn_ocr_b200 as R
        from rcnn_ocr_b200.dist import GradAllReducer
        self.R = R
        torch.manual_seed(0)
        self.model = R.RCNN(194, hidden_size=hidden).to(device).to(memory_format=torch.channels_last)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=5.1e-4, weight_decay=1.95e-5, fused=True, capturable=True)
        self.reducer = GradAllReducer(self.params) if world > 1 else None

    def __call__(self, x, tg, il, tl):
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feats = self.model._features(x)
        logits = self.model.ctc_head(self.model._encode_features(feats))
        loss = self.R.ctc_loss_from_logits(logits.permute(1, 0, 2), tg, il, tl, 0, "mean", True,
                                           max_target_length=tg.shape[1])
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss


def run_cfg4(args, device, world, rank, timed_fn, hidden=256):
    from rcnn_ocr_b200.dist import shard_range
    import rcnn_ocr_b200 as R
    lo, hi = shard_range(FullTrainStep.GLOBAL_B, rank, world)
    Bl = hi - lo
    step = FullTrainStep(device, world, hidden)
    with torch.no_grad():
        T = int(step.model._features(torch.zeros(1, 3, 32, 128, device=device)).shape[1])
    lmax = max(1, min(32, T // 2))            # labels U{1..lmax}: a CTC alignment must fit T frames
    batches = []
    for i in range(4):
        g = torch.Generator().manual_seed(4321 + 31 * rank + i)
        x = (torch.rand(Bl, 3, 32, 128, generator=g) * 2 - 1).to(device).contiguous(memory_format=torch.channels_last)
        tl = torch.randint(1, lmax + 1, (Bl,), generator=g).to(device)
        tg = torch.randint(4, 195, (Bl, lmax), generator=g).to(device)      # itos[3:] -> CTC classes 4..194
        batches.append([x, tg, torch.full((Bl,), T, dtype=torch.long, device=device), tl])
    fn, mode = step, "eager"
    # cuDNN picks the backbone's forward / dgrad / wgrad algorithms by measurement during the warm-up steps (before the capture)
    bench_prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = os.environ.get("RCNN_CUDNN_BENCHMARK", "1") == "1"
    if args.graph and (world == 1 or args.graph_dp):
        try:
            fn = R.GraphedStep(step, batches[0])
            mode = "cuda-graph replay"
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            fn, mode = step, f"eager (capture failed: {type(e).__name__})"
    ms = timed_fn(lambda i: fn(*batches[i % 4]))
    torch.backends.cudnn.benchmark = bench_prev
    # share of the step that is this repo's kernels: the same step without the backbone (features as inputs)
    return {"ms": ms, "mode": mode, "T": T, "lmax": lmax, "per_gpu_batch": Bl, "hidden": hidden, "keep": (step, fn)}



def run_cfg2(args, device, timed_fn, peaks, cpu: bool):
    """BASELINE configs[1]: CTC loss fwd+bwd microbench, T=64, C=195, labels U{1..32}, batch 256, against torch's CTCLoss
    (SURVEY 8d cfg 2).  Ours: ONE launch from the logits (log_softmax + loss + gradient at the logits); torch: log_softmax,
    ctc_loss, ctc_loss_backward, log_softmax_backward (ATen CUDA on the same GPU; torch CPU on the host cores)."""
    import torch.nn.functional as F
    import rcnn_ocr_b200 as R
    from rcnn_ocr_b200 import _lib
    T, C, N = CFG["T"], CFG["C"], CFG["B"]
    g = torch.Generator().manual_seed(1234)
    xs = [torch.randn(N, T, C, generator=g).to(device).requires_grad_(True) for _ in range(16)]   # 16 x 12.8 MB x (x + grad) > L2
    tl = torch.randint(1, 33, (N,), generator=g)
    tg = torch.randint(1, C, (N, 32), generator=g)
    il = torch.full((N,), T)
    tgd, ild, tld = tg.to(device), il.to(device), tl.to(device)

    def ours(i):
        x = xs[i % 16]
        x.grad = None
        R.ctc_loss_from_logits(x.permute(1, 0, 2), tgd, ild, tld, 0, "mean", True, max_target_length=32).backward()

    def aten(i):
        x = xs[i % 16]
        x.grad = None
        F.ctc_loss(F.log_softmax(x, 2).permute(1, 0, 2), tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()

    ms_api = timed_fn(ours)
    _lib.prof_enable(True)
    for i in range(args.steps):
        ours(i)
    torch.cuda.synchronize()
    kms, n = _lib.prof_read(1)
    _lib.prof_enable(False)
    _lib.lib().rcnn_prof_reset()
    ms_k = kms / max(n, 1)
    ms_aten = timed_fn(aten)
    alg = 2.0 * T * C * 4 * N
    out = {"ours": {"kernel_us": round(ms_k * 1e3, 2), "api_fwd_bwd_us": round(ms_api * 1e3, 2),
                    "GBps": round(alg / (ms_k * 1e-3) / 1e9, 1), "hbm_frac": round(alg / (ms_k * 1e-3) / 1e9 / peaks["hbm"], 4),
                    "seq_per_s": round(N / (ms_k * 1e-3))},
           "torch_cuda": {"api_fwd_bwd_us": round(ms_aten * 1e3, 2), "seq_per_s": round(N / (ms_aten * 1e-3)),
                          "what": "F.log_softmax + F.ctc_loss + backward, ATen CUDA, eager, same GPU and inputs"},
           "speedup_vs_torch_cuda": round(ms_aten / ms_api, 2),
           "config": {"T": T, "N": N, "C": C, "label_len": "U{1..32}", "reduction": "mean", "zero_infinity": True,
                      "algorithmic_bytes_per_seq": 2 * T * C * 4}}
    if cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        xc = xs[0].detach().cpu().requires_grad_(True)
        best = None
        for _ in range(4):
            xc.grad = None
            t0 = time.perf_counter()
            F.ctc_loss(F.log_softmax(xc, 2).permute(1, 0, 2), tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["torch_cpu"] = {"fwd_bwd_ms": round(best * 1e3, 2), "seq_per_s": round(N / best), "cores": os.cpu_count(),
                            "what": "the same torch ops on the host cores (min of 3 after a warm-up)"}
    return out


def run_b512(args, step, infer, device, timed_fn, sync, peaks, use_graph):
    """cfg B at 512 lines per GPU (BASELINE configs[3] names a batch of 512): 16 (direction, 64-sequence) work items for
    8 CTA groups.  The backward recurrent kernel then works on two items per group at a time (rcnn_lstm_plan)."""
    from rcnn_ocr_b200 import _lib, ops
    B2, T, H = 2 * CFG["B"], CFG["T"], CFG["H"]
    ring = [[t.to(device) for t in make_batch(B2, 4321 + i)] for i in range(4)]
    ring = [[b[0].to(torch.bfloat16)] + b[1:] for b in ring]
    step(*ring[0])
    sync()
    g = step.R.GraphedStep(step, ring[0]) if use_graph else step
    ms = timed_fn(lambda i: g(*ring[i % 4]))
    step.enc.eval(); step.head.eval()
    gi = step.R.GraphedStep(infer.device, [ring[0][0]]) if use_graph else infer.device
    ms_inf = timed_fn(lambda i: gi(ring[i % 4][0]))
    step.enc.train(); step.head.train()
    _lib.prof_enable(True)
    for i in range(args.steps):
        step(*ring[i % 4])
    sync()
    kern = {}
    rec_flops = 2.0 * B2 * T * H * 4 * H * 2
    for kid, name in {3: "lstm_fwd", 4: "lstm_bwd"}.items():
        kms, n = _lib.prof_read(kid)
        if n:
            per = kms / n
            fl = rec_flops * (2 if name == "lstm_fwd" else 1)     # the forward also multiplies W_ih x_t
            kern[name] = {"ms_per_launch": round(per, 4), "us_per_timestep": round(per * 1e3 / T, 3),
                          "tensor_frac": round(fl / (per * 1e-3) / 1e12 / peaks["tf_sus"], 4)}
    _lib.prof_enable(False)
    _lib.lib().rcnn_prof_reset()
    del g, gi
    pf, pb = ops.lstm_plan(B2, H), ops.lstm_plan(B2, H, backward=True)
    return {"value": round(B2 / (ms * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms, 4), "per_gpu_batch": B2,
            "infer": {"value": round(B2 / (ms_inf * 1e-3), 1), "ms_per_step": round(ms_inf, 4)},
            "kernels": kern,
            "plan": {"forward": {"items_per_group": pf[0], "groups": pf[1]}, "backward": {"items_per_group": pb[0], "groups": pb[1]}},
            "note": "same train / inference step as the headline at 512 lines per GPU, device-resident, "
                    + ("CUDA-graph replay" if use_graph else "eager")}


def timed(fn, steps, warmup, sync, barrier):
    for i in range(warmup):
        fn(i)
    barrier()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    sync()
    barrier()
    return e0.elapsed_time(e1) / steps


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from rcnn_ocr_b200 import _lib, ops
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        # NCCL's all-reduce kernels run beside this library's kernels on the SMs those leave free (the recurrent kernels
        # 20, the GEMMs 16 under data parallelism: dist.GradAllReducer): cap their CTA count accordingly
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=device)
    _lib.check(_lib.lib().rcnn_device_check(), "rcnn_device_check")
    peaks = load_peaks()
    B = args.batch
    RING = 8

    def barrier():
        if world > 1:
            dist.barrier()

    def sync():
        torch.cuda.synchronize(device)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the feature columns arrive as bf16: the dtype the autocast backbone emits (training/train.py:499) and the one
    # BASELINE configs[2] names; the CPU arm keeps fp32
    host = []
    for i in range(RING):
        f, tg_, il_, tl_ = make_batch(B, 1234 + 17 * rank + i)
        host.append([t.pin_memory() for t in (f.to(torch.bfloat16), tg_, il_, tl_)])
    dev = [[t.to(device) for t in b] for b in host]
    step = TrainStep(device, world, rank)
    infer = InferStep(step)
    use_graph = args.graph and (world == 1 or args.graph_dp)
    # One CUDA-graph replay per step (rcnn_ocr_b200.GraphedStep): the step function is the same eager
    # code; inputs are copied into the graph's static tensors inside the timed region.
    gstep = step.R.GraphedStep(step, dev[0]) if use_graph else step

    # ---- device-resident train step (headline `value`) -----------------------------------------
    l0 = ops.launch_count()
    step(*dev[0])                                        # one eager step: counts this library's launches per step
    sync()
    launches = ops.launch_count() - l0                   # (a graph replay runs exactly the kernels captured from it)
    # two captured graphs used in turn; the next step's inputs are copied into the idle graph's static tensors on a copy
    # stream while the other graph runs (device-resident loop: from the ring in HBM; end-to-end loop: from pinned host
    # memory).  At N > 1 both graphs hold the NCCL bucket all-reduces and every rank replays them in the same order.
    copy_stream = torch.cuda.Stream()
    pingpong = use_graph
    gsteps = [gstep, step.R.GraphedStep(step, dev[0])] if pingpong else None
    slots = [g.static_in for g in gsteps] if pingpong else \
        [[torch.empty_like(t, device=device) for t in host[0]] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def stage(i, source=None):
        s = i % 2
        src = (source or host)[i % RING]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            for d, h in zip(slots[s], src):
                d.copy_(h, non_blocking=True)
            ready[s].record(copy_stream)

    def dev_train(i):                                   # inputs resident in HBM (ring of 8 sets > L2)
        if not pingpong:
            return gstep(*dev[i % RING])
        s = i % 2
        if dev_train.primed is False:
            stage(i, dev)
            dev_train.primed = True
        stage(i + 1, dev)
        torch.cuda.current_stream().wait_event(ready[s])
        out = gsteps[s].replay()
        freed[s].record()
        return out

    dev_train.primed = False
    for f in freed:
        f.record()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_train = timed(dev_train, args.steps, args.warmup, sync, barrier)
    clk = clocks.stop()
    ms_train = max_over_ranks(ms_train)
    sync()

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the loss, every step ----------

    losses = []

    def e2e_train(i):
        s = i % 2
        if e2e_train.primed is False:
            stage(i)
            e2e_train.primed = True
        stage(i + 1)                                    # overlap the next step's H2D with this step
        torch.cuda.current_stream().wait_event(ready[s])
        loss = gsteps[s].replay() if pingpong else gstep(*slots[s])   # (one graph: d2d copy of the staged inputs + replay)
        freed[s].record()
        losses.append(loss.item())                      # D2H read of the result, every step

    e2e_train.primed = False
    for f in freed:
        f.record()
    ms_e2e = max_over_ranks(timed(e2e_train, args.steps, args.warmup, sync, barrier))
    sync()
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- greedy inference (same encoder, decode on device); eval() as inference.py:88 does -------
    step.enc.eval(); step.head.eval()
    ginfer = step.R.GraphedStep(infer.device, [dev[0][0]]) if use_graph else infer.device
    ms_inf = max_over_ranks(timed(lambda i: ginfer(dev[i % RING][0]), args.steps, args.warmup, sync, barrier))

    def stage_inf(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            slots[s][0].copy_(host[i % RING][0], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_infer(i):                                    # pinned host features in, python strings out
        s = i % 2
        if e2e_infer.primed is False:
            stage_inf(i)
            e2e_infer.primed = True
        stage_inf(i + 1)                                 # next batch's H2D overlaps this batch's kernels
        torch.cuda.current_stream().wait_event(ready[s])
        ids, lens = ginfer(slots[s][0])
        freed[s].record()
        # serving loop: the D2H of this batch's ids / lens is enqueued behind its kernels, the strings of the PREVIOUS batch
        # are built while this batch runs (every timed step still copies one batch in, one out and maps one to strings)
        pending, e2e_infer.pending = e2e_infer.pending, step.R.ids_to_text_async(ids, lens, infer.alphabet)
        return pending.result()[0] if pending is not None else None

    e2e_infer.primed = False
    e2e_infer.pending = None
    for f in freed:
        f.record()
    ms_inf_e2e = max_over_ranks(timed(e2e_infer, args.steps, args.warmup, sync, barrier))
    # ---- validation metrics of one batch on the device (K5): decode ids vs the CTC targets ---------------
    table = step.R.CharsetTable(infer.alphabet, device)
    v_ids, v_lens = [t.clone() for t in ginfer(dev[0][0])]
    sync()
    t0 = time.perf_counter()
    for _ in range(10):
        val = step.R.validation_metrics(v_ids, v_lens, dev[0][1], dev[0][3], table)
    ms_val = (time.perf_counter() - t0) / 10 * 1e3
    # ---- attention decoder (section 8f-1, the reference's live decoder): greedy decode of one batch ------
    attn_res = None
    if not args.no_attention:
        torch.manual_seed(1)
        attn = step.R.Attention(CFG["H"], CFG["H"], CFG["C"] - 1, 1, 2, 0, 3).to(device).eval()
        encs = [torch.randn(B, CFG["T"], CFG["H"], device=device) for _ in range(RING)]   # (fp32: the reference's encoder output)
        decode = lambda e: attn(e, is_train=False, batch_max_length=25)
        gdec = step.R.GraphedStep(decode, [encs[0]]) if use_graph else decode
        ms_attn = max_over_ranks(timed(lambda i: gdec(encs[i % RING]), args.steps, args.warmup, sync, barrier))
        attn_res = {"value": round(B * world / (ms_attn * 1e-3), 1), "unit": "lines/s", "ms_per_batch": round(ms_attn, 4),
                    "config": {"B": B, "T_enc": CFG["T"], "hidden": CFG["H"], "classes": CFG["C"] - 1, "steps": 26},
                    "note": "Attention._greedy_decode (model/model.py:89-108) on the device: hoisted i2h GEMM + 26 x "
                            "(score/softmax/context with the previous step's mask+argmax, gate GEMM with the LSTMCell step in "
                            "its epilogue, one GEMM for h2h | generator), launched as a programmatic-dependent chain"}
        # the decoder's training step (the reference's live loss path, training/train.py:499-505): teacher forcing with the
        # default alpha dropout 0.1, cross entropy, backward, Adam -- fused forward / backward kernels, one graph replay
        import torch.nn.functional as F
        attn_t = step.R.Attention(CFG["H"], CFG["H"], CFG["C"] - 1, 1, 2, 0, 3, dropout_p=0.1).to(device).train()
        opt_t = torch.optim.Adam(attn_t.parameters(), lr=5.1e-4, fused=True, capturable=True)
        Vt, St = CFG["C"] - 1, 26
        gt = torch.Generator().manual_seed(5 + rank)
        texts = [torch.randint(4, Vt, (B, St + 1), generator=gt).to(device) for _ in range(RING)]
        for tx in texts:
            tx[:, 0] = 1

        def attn_train(e, tx):
            opt_t.zero_grad(set_to_none=True)
            logits = attn_t(e.detach().requires_grad_(True), tx[:, :St], is_train=True, batch_max_length=St - 1)
            loss = F.cross_entropy(logits.reshape(-1, Vt), tx[:, 1:].reshape(-1), ignore_index=0)
            loss.backward()
            opt_t.step()
            return loss

        ms_at_eager = max_over_ranks(timed(lambda i: attn_train(encs[i % RING], texts[i % RING]), min(args.steps, 10), 3, sync, barrier))
        g_at = step.R.GraphedStep(attn_train, [encs[0], texts[0]]) if use_graph else attn_train
        ms_at = max_over_ranks(timed(lambda i: g_at(encs[i % RING], texts[i % RING]), args.steps, args.warmup, sync, barrier))
        attn_res["train"] = {"value": round(B * world / (ms_at * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms_at, 4),
                             "ms_per_step_eager_launches": round(ms_at_eager, 4),
                             "note": "teacher-forced forward (26 steps, alpha dropout 0.1) + cross entropy + backward + Adam on "
                                     "the fused kernels (attention._TeacherForcedFn), one CUDA-graph replay per step"}
        if world > 1:
            attn_res["train"]["note"] += "; at N > 1: independent per-GPU replicas of the decoder (no gradient all-reduce in this sub-measurement)"
        if rank == 0 and world == 1 and not args.no_extras:
            from oracle.ref_port import RefAttention
            ref_t = RefAttention(CFG["H"], CFG["H"], Vt).to(device).train()
            ref_t.load_state_dict(attn_t.state_dict(), strict=True)
            opt_r = torch.optim.Adam(ref_t.parameters(), lr=5.1e-4, fused=True)

            def vendor_train(e, tx):                      # model/model.py:110-148 op by op on torch's CUDA kernels, bf16 autocast
                opt_r.zero_grad(set_to_none=True)
                e = e.detach().requires_grad_(True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    cell = ref_t.attention_cell
                    h = e.new_zeros(B, CFG["H"]); c = e.new_zeros(B, CFG["H"])
                    hs = []
                    for t in range(St):
                        onehot = F.one_hot(tx[:, t], Vt).float()
                        ee = cell.score(torch.tanh(cell.i2h(e) + cell.h2h(h).unsqueeze(1)))
                        alpha = F.dropout(F.softmax(ee, dim=1), p=0.1, training=True)
                        context = torch.bmm(alpha.transpose(1, 2), e).squeeze(1)
                        h, c = cell.rnn(torch.cat([context, onehot], 1), (h, c))
                        hs.append(h)
                    logits = ref_t.generator(torch.stack(hs, 1)).float()
                loss = F.cross_entropy(logits.reshape(-1, Vt), tx[:, 1:].reshape(-1), ignore_index=0)
                loss.backward()
                opt_r.step()
                return loss

            ms_vt = timed(lambda i: vendor_train(encs[i % RING], texts[i % RING]), 5, 3, sync, lambda: None)
            attn_res["train"]["vendor"] = {"value": round(B / (ms_vt * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms_vt, 4),
                                           "note": "the reference's op sequence (nn.Linear / nn.LSTMCell / bmm, proj_H recomputed "
                                                   "every step) on torch's CUDA kernels under bf16 autocast, eager, same GPU"}
            attn_res["train"]["vs_vendor"] = round(ms_vt / ms_at, 2)
            del ref_t, opt_r
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle.ref_port import RefAttention
            torch.set_num_threads(os.cpu_count() or 1)
            ref = RefAttention(CFG["H"], CFG["H"], CFG["C"] - 1).eval()
            xb = torch.randn(32, CFG["T"], CFG["H"])
            ref.greedy(xb[:4])
            t0 = time.perf_counter()
            ref.greedy(xb)
            dt = time.perf_counter() - t0
            attn_res["cpu_baseline"] = {"value": round(32 / dt, 1), "unit": "lines/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "32 lines x 26 steps, oracle/ref_port.RefAttention (the reference's op sequence)"}
    # ---- cfg 5: inference sweep over the batch size (global batch B5, sharded over the ranks, no collective) ----
    from rcnn_ocr_b200.dist import shard_range
    tf = lambda fn: max_over_ranks(timed(fn, args.steps, args.warmup, sync, barrier))
    sweep = []
    for B5 in ([] if args.no_extras else [1, 32, 256, 4096]):
        lo5, hi5 = shard_range(B5, rank, world)
        if B5 < world:
            continue
        g5 = torch.Generator().manual_seed(99 + rank)
        f5 = [torch.randn(hi5 - lo5, CFG["T"], CFG["IN"], generator=g5).to(torch.bfloat16).to(device)
              for _ in range(2 if B5 >= 4096 else 4)]
        fn5 = step.R.GraphedStep(infer.device, [f5[0]]) if use_graph else infer.device
        ms5 = tf(lambda i: fn5(f5[i % len(f5)]))
        sweep.append({"global_batch": B5, "per_gpu_batch": hi5 - lo5, "ms_per_batch": round(ms5, 4),
                      "value": round(B5 / (ms5 * 1e-3), 1), "unit": "lines/s"})
        del fn5, f5
    step.enc.train(); step.head.train()

    # ---- vendor bar, cfg 1, cfg 4 (strong scaling) ------------------------------------------------------------
    vendor = cfg1 = cfg2 = cfg4 = b512 = None
    if not args.no_extras:
        if world == 1:
            b512 = run_b512(args, step, infer, device, tf, sync, peaks, use_graph)
            cfg2 = run_cfg2(args, device, tf, peaks, cpu=(rank == 0 and not args.no_cpu_baseline))
            vendor = run_vendor(args, dev, B, tf)
            cfg1 = run_cfg1(args, device, tf, cpu=(rank == 0 and not args.no_cpu_baseline))
        r4 = run_cfg4(args, device, world, rank, tf, hidden=256)
        gb = FullTrainStep.GLOBAL_B
        cfg4 = {"value": round(gb / (r4["ms"] * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(r4["ms"], 4),
                "scaling": "strong", "global_batch": gb, "per_gpu_batch": r4["per_gpu_batch"], "launch_mode": r4["mode"],
                "config": {"workload": "cfg4 full train step (training/train.py:493-508, configs/config.json:20-29): images "
                                       "[B,3,32,128] -> SE-ResNet31 (torch/cuDNN, channels_last, bf16 autocast) -> enc_rnn "
                                       f"(hidden {r4['hidden']}) -> CTC head -> fused CTC loss -> backward -> all-reduce of all "
                                       "gradients -> Adam", "T": r4["T"], "label_len": f"U{{1..{r4['lmax']}}}",
                           "parallelism": f"dp{world}"},
                "limiter": "the cuDNN backbone (98 % of the step's FLOPs, SURVEY 8d): at 512 / N lines per GPU its convolutions on "
                           "32 x 128 images stop filling the GPU (27 ms at 512 lines, 6.6 ms at 64), and 184 MB of fp32 gradients are "
                           "all-reduced per step; the hand-written part (enc_rnn H = 256, T = 16 + CTC) is < 1 ms of it"}
        del r4

    # ---- per-kernel timing for the roofline (separate pass; CUDA events on the launching stream) -
    _lib.prof_enable(True)
    for i in range(args.steps):
        step(*dev[i % RING])
    sync()
    names = {0: "decode", 1: "ctc", 2: "gemm_tn", 3: "lstm_fwd", 4: "lstm_bwd", 5: "gemm_atb"}
    kern = {}
    for kid, name in names.items():
        ms, n = _lib.prof_read(kid)
        if n:
            kern[name] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps}
    _lib.lib().rcnn_prof_reset()
    step.enc.eval(); step.head.eval()
    for i in range(args.steps):
        infer.device(dev[i % RING][0])
    sync()
    step.enc.train(); step.head.train()
    ms_dec, n_dec = _lib.prof_read(0)
    _lib.prof_enable(False)

    T, H, C, IN = CFG["T"], CFG["H"], CFG["C"], CFG["IN"]
    rec_flops = 2.0 * B * T * H * 4 * H * 2                 # recurrent matmuls of one block, both directions
    BT = 2.0 * B * T
    # forward recurrence launches also multiply W_ih x_t (input projection fused into the kernel)
    fwd_flops = [rec_flops + 2.0 * B * T * IN * 4 * H * 2, rec_flops + 2.0 * B * T * H * 4 * H * 2]   # block 1, block 2
    gemm_flops = {   # per train step (block 1 has no dX: its input needs no gradient)
        "gemm_tn": BT * (2 * H * H + 2 * H * H + H * C)                               # forward: linear (x2 blocks), head
                   + BT * (2 * (H * 2 * H) + H * 8 * H + C * H),                      # backward: dhcat (x2), dX of block 2, d enc
        "gemm_atb": BT * (IN * 8 * H + H * 8 * H + 2 * (H * 8 * H) + 2 * (2 * H * H) + C * H),   # dW_ih, dW_hh, dW_lin, dW_head
    }
    roof = None
    if kern:
        dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
        k = kern[dom]
        per_launch_ms = k["ms_per_step"] / k["launches_per_step"]
        if dom in ("lstm_fwd", "lstm_bwd"):
            launch_flops = sum(fwd_flops) / 2 if dom == "lstm_fwd" else rec_flops
            ach = launch_flops / (per_launch_ms * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tf_sus"],
                    "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sus"], 4), "traffic": load_traffic(dom),
                    "traffic_note": "DRAM bytes per launch, ncu --set full capture at this config (profiles/ncu_traffic_r02.json)",
                    "algorithmic": f"2*B*T*H*4H*2dirs = {rec_flops / 1e9:.1f} GFLOP of recurrent matmuls per launch (one block)"
                                   + (f" + {(launch_flops - rec_flops) / 1e9:.1f} GFLOP of fused input projection" if dom == "lstm_fwd" else ""),
                    "us_per_timestep": round(per_launch_ms * 1e3 / T, 3), "peak_source": peaks["src"]}
        elif dom in ("gemm_tn", "gemm_atb"):
            ach = gemm_flops[dom] / (k["ms_per_step"] * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tf_sus"],
                    "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sus"], 4), "traffic": None,
                    "algorithmic": f"{gemm_flops[dom] / 1e9:.1f} GFLOP over the {dom} launches of a step",
                    "peak_source": peaks["src"]}
        else:
            nbytes = 2.0 * T * C * 4 * B
            ach = nbytes / (per_launch_ms * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4), "traffic": None, "peak_source": peaks["src"]}
    kernels = {n: {"ms_per_step": round(v["ms_per_step"], 4), "launches_per_step": v["launches_per_step"]}
               for n, v in kern.items()}
    for n in ("gemm_tn", "gemm_atb"):
        if n in kern:
            kernels[n]["tflops"] = round(gemm_flops[n] / (kern[n]["ms_per_step"] * 1e-3) / 1e12, 1)
            kernels[n]["tensor_frac"] = round(kernels[n]["tflops"] / peaks["tf_sus"], 4)
    for n in ("lstm_fwd", "lstm_bwd"):
        if n in kern:
            per = kern[n]["ms_per_step"] / kern[n]["launches_per_step"]
            kernels[n]["us_per_timestep"] = round(per * 1e3 / T, 3)
            fl = sum(fwd_flops) / 2 if n == "lstm_fwd" else rec_flops
            kernels[n]["tensor_frac"] = round(fl / (per * 1e-3) / 1e12 / peaks["tf_sus"], 4)
    if "ctc" in kern:
        per = kern["ctc"]["ms_per_step"] / kern["ctc"]["launches_per_step"]
        kernels["ctc"]["hbm_frac"] = round(2.0 * T * C * 4 * B / (per * 1e-3) / 1e9 / peaks["hbm"], 4)
    if n_dec:
        per = ms_dec / n_dec
        kernels["decode"] = {"ms_per_launch": round(per, 4),
                             "hbm_frac": round((T * C * 4 + T * 4 + 4) * B / (per * 1e-3) / 1e9 / peaks["hbm"], 4)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = cpu_reference(2, 1, B)
            cpu = {"value": round(v, 2), "unit": "lines/s", "cores": cores, "kind": "port",
                   "sample": f"{B} lines/step x 2 steps (after 1 warm-up) of the same train-step workload on the "
                             "host CPU: oracle/ref_port.py = the reference's torch-CPU op sequence "
                             "(nn.LSTM+nn.Linear x2, log_softmax+F.ctc_loss, backward)",
                   "ms_per_step": round(ms, 1)}
        total_B = B * world
        line = {
            "metric": "text-lines/sec", "value": round(total_B / (ms_train * 1e-3), 1), "unit": "lines/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_train, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, B, world),
            "e2e": {"value": round(total_B / (ms_e2e * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "note": "pinned host features/targets -> H2D (double-buffered on a copy stream"
                            + (", straight into the static inputs of two captured graphs used in turn" if pingpong else "")
                            + ") -> train step -> loss.item() every step"},
            "infer": {"value": round(total_B / (ms_inf * 1e-3), 1), "unit": "lines/s", "ms_per_step": round(ms_inf, 4),
                      "e2e": {"value": round(total_B / (ms_inf_e2e * 1e-3), 1), "ms_per_step": round(ms_inf_e2e, 4),
                              "h2d_bytes_per_step": host[0][0].numel() * host[0][0].element_size(), "d2h_bytes_per_step": B * (T + 1) * 4,
                              "note": "pinned host features -> H2D (double-buffered on a copy stream) -> encoder + "
                                      "greedy decode -> D2H of ids/len -> python strings (the strings of batch i are built "
                                      "while batch i+1 runs: ids_to_text_async)"},
                      "val_metrics": {"ms_per_batch": round(ms_val, 3), "cer": round(val["cer"], 4),
                                      "accuracy": val["accuracy"],
                                      "note": "CER/WER/accuracy of one decoded batch on the device (K5 edit distance), "
                                              "wall clock incl. the D2H of the per-pair integers"},
                      "attention_decoder": attn_res},
            "launch_mode": ("cuda-graph replay (one graph = the whole step; two graphs used in turn, the next step's inputs are "
                            "copied from the ring in HBM into the idle graph's static tensors on a copy stream)") if use_graph else "eager",
            "gpu_launches": round(launches * args.steps),
            "gpu_launches_per_step": round(launches, 1),
            "vendor": vendor,
            "vs_vendor": ({"train": round(total_B / (ms_train * 1e-3) / vendor["value"], 3),
                           "infer": round(total_B / (ms_inf * 1e-3) / vendor["infer"]["value"], 3)} if vendor else None),
            "cfgB_batch512": b512, "cfg2_ctc_microbench": cfg2,
            "cfg1_minimal_inference": cfg1, "cfg4_strong": cfg4, "cfg5_infer_sweep": sweep or None,
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "clocks": clk,
            "loss_first_last": [round(losses[0], 4), round(losses[-1], 4)] if losses else None,
        }
        emit(line)
    if world > 1:
        # Captured graphs hold NCCL work: drop them before the process group; a watchdog ends the process
        # if the teardown still blocks (the result line is already out).
        import gc
        threading.Timer(20.0, lambda: os._exit(0)).start()
        del gstep, ginfer
        gc.collect()
        sync()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


_RESULT_OUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's original stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the result line only: file descriptor 1 is pointed at stderr for everything else that
    # prints from native code (NCCL's "NCCL version ..." banner under NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE)
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=CFG["B"], help="lines per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=64, help="lines per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-attention", action="store_true", help="skip the attention-decoder (section 8f-1) measurement")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the vendor bar, cfg 1 (minimal_inference), cfg 4 (strong-scaled full step) and the cfg 5 sweep")
    ap.add_argument("--no-graph-dp", dest="graph_dp", action="store_false",
                    help="N>1: launch eagerly (default: the NCCL bucket all-reduces on the side stream join the capture)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel eagerly instead of replaying the captured CUDA graph (N=1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        emit({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"})
        sys.exit(2)
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
