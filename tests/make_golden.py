"""Generate tests/golden/*.npz from the REFERENCE's own code and torch float64.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box):   python tests/make_golden.py

  decode_*.npz   training/utils.py:122-150  ctc_greedy_decoder  (reference code, imported)
  decodefn_*.npz training/utils.py:153-162  decode              (reference code, imported)
  cnn_*.npz      model/seresnet31.py:70-187 SEResNet31          (reference code, imported)
  bilstm_*.npz   model/model.py:151-163     BidirectionalLSTM   (reference code, imported)
  encrnn_*.npz   model/model.py:195-198     Sequential of two blocks, outputs + all grads
  attn_*.npz     model/model.py:50-148      Attention (reference code, imported), eval mode
  ctc_*.npz      torch.nn.functional.ctc_loss, CPU float64 (the reference has no CTC code;
                 SURVEY.md section 0) -- loss per reduction and the gradient AT THE LOGITS.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
sys.path.insert(0, REF)
from model.model import BidirectionalLSTM  # noqa: E402
from training.utils import ctc_greedy_decoder, decode as ref_decode  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
os.makedirs(OUT, exist_ok=True)


def charset():
    itos = []
    with open(os.path.join(REF, "configs", "charset.txt"), encoding="utf-8") as fh:
        for line in fh:
            tok = line.rstrip("\n")
            if tok:
                itos.append(tok)
    return itos


def gen_decode():
    itos = charset()
    g = torch.Generator().manual_seed(1234)
    cases = {}
    # random, B >= T (no permute), full charset
    cases["rand_b8_t6"] = (torch.randn(8, 6, len(itos) + 1, generator=g), itos)
    # B < T in batch-first layout: the reference heuristic permutes (utils.py:132-133)
    cases["heur_b3_t7"] = (torch.randn(3, 7, len(itos) + 1, generator=g), itos)
    # time-major input with T < B: heuristic permutes to batch-first
    cases["tmajor_t5_b9"] = (torch.randn(5, 9, 11, generator=g), list("abcdefghij"))
    # ties: quantised logits -> many equal maxima (first index must win)
    cases["ties"] = (torch.randint(0, 3, (16, 12, 7), generator=g).float(), list("abcdef"))
    # all blank / all same label / alternating
    x = torch.full((4, 4, 5), -1.0)
    x[0, :, 0] = 1.0
    x[1, :, 2] = 1.0
    x[2, 0::2, 3] = 1.0
    x[2, 1::2, 0] = 1.0
    x[3, 0, 1] = 1.0; x[3, 1, 1] = 1.0; x[3, 2, 0] = 1.0; x[3, 3, 1] = 1.0
    cases["patterns"] = (x, list("wxyz"))
    # NaN and inf entries
    y = torch.randn(6, 5, 9, generator=g)
    y[0, 0, 4] = float("nan"); y[1, 2, 0] = float("nan"); y[1, 2, 7] = float("nan")
    y[2, 1, 3] = float("inf"); y[3, :, :] = float("-inf"); y[4, 3, 8] = float("inf"); y[4, 3, 2] = float("inf")
    cases["nan_inf"] = (y, list("abcdefgh"))
    # cfg-B-like shape, reduced batch
    cases["cfgb_b16_t16"] = (torch.randn(16, 16, len(itos) + 1, generator=g), itos)
    for name, (logits, alpha) in cases.items():
        texts, seqs = ctc_greedy_decoder(logits, alpha, blank=0)
        np.savez_compressed(os.path.join(OUT, f"decode_{name}.npz"), logits=logits.numpy(),
                            alphabet=json.dumps(alpha), texts=json.dumps(texts), seqs=json.dumps(seqs))
    # non-zero blank
    logits = torch.randn(10, 8, 6, generator=g)
    texts, seqs = ctc_greedy_decoder(logits, list("abcdef"), blank=3)
    np.savez_compressed(os.path.join(OUT, "decode_blank3.npz"), logits=logits.numpy(), blank=3,
                        alphabet=json.dumps(list("abcdef")), texts=json.dumps(texts), seqs=json.dumps(seqs))


def gen_decode_fn():
    """decodefn_*.npz: the reference's ``decode(ctc_out, alphabet, method)`` (training/utils.py:153-162) --
    tuple unwrap, log_softmax, greedy -- called exactly as a validation loop would call it."""
    itos = charset()
    g = torch.Generator().manual_seed(4321)
    cases = {
        "tuple_b12_t9": ((torch.randn(12, 9, len(itos) + 1, generator=g) * 3, torch.zeros(1)), itos),   # tuple input
        "plain_b6_t20": (torch.randn(6, 20, 11, generator=g), list("abcdefghij")),                     # B < T: heuristic
        "cfga_b32_t16": ((torch.randn(32, 16, len(itos) + 1, generator=g),), itos),                    # config.json geometry
    }
    for name, (out, alpha) in cases.items():
        texts, seqs = ref_decode(out, alpha, method="greedy")
        logits = out[0] if isinstance(out, tuple) else out
        np.savez_compressed(os.path.join(OUT, f"decodefn_{name}.npz"), logits=logits.numpy(),
                            is_tuple=np.array(int(isinstance(out, tuple))), alphabet=json.dumps(alpha),
                            texts=json.dumps(texts), seqs=json.dumps(seqs))


def gen_cnn():
    """cnn_*.npz: the reference's SEResNet31 (model/seresnet31.py:70-187, imported) in eval mode on recipe weights
    (conftest.recipe_state_dict: a function of the state-dict keys only), input and output stored."""
    from model.seresnet31 import SEResNet31
    from conftest import recipe_state_dict
    ref = SEResNet31(3, 512).eval()
    ref.load_state_dict(recipe_state_dict(ref, 77), strict=True)
    g = torch.Generator().manual_seed(78)
    x = torch.rand(2, 3, 32, 64, generator=g) * 2 - 1
    with torch.no_grad():
        y = ref(x)
    np.savez_compressed(os.path.join(OUT, "cnn_eval_2x32x64.npz"), x=x.numpy(), y=y.numpy(), seed=np.array(77),
                        keys=json.dumps(sorted(ref.state_dict().keys())))
    # training mode (batch statistics), one forward + backward: output, input gradient, one weight gradient
    ref.train()
    ref.load_state_dict(recipe_state_dict(ref, 79), strict=True)
    xt = (torch.rand(4, 3, 32, 64, generator=g) * 2 - 1).requires_grad_(True)
    yt = ref(xt)
    yt.square().mean().backward()
    np.savez_compressed(os.path.join(OUT, "cnn_train_4x32x64.npz"), x=xt.detach().numpy(), y=yt.detach().numpy(),
                        dx=xt.grad.numpy(), dw0=ref.conv0[0].weight.grad.numpy(), seed=np.array(79))


def gen_bilstm():
    for name, (B, T, I, H, O) in {"tiny": (3, 5, 8, 4, 6), "odd": (2, 7, 24, 16, 8),
                                   "h32": (4, 9, 32, 32, 32)}.items():
        torch.manual_seed(7)
        m = BidirectionalLSTM(I, H, O).double()
        x = torch.randn(B, T, I, dtype=torch.float64, requires_grad=True)
        # non-contiguous batch-first view, as produced by model/model.py:218
        xin = x.permute(0, 2, 1).contiguous().permute(0, 2, 1)
        out = m(xin)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        d = {f"p.{k}": v.detach().numpy() for k, v in m.state_dict().items()}
        d.update({f"g.{k}": v.grad.numpy() for k, v in m.named_parameters()})
        np.savez_compressed(os.path.join(OUT, f"bilstm_{name}.npz"), x=x.detach().numpy(),
                            out=out.detach().numpy(), w=w.numpy(), gx=x.grad.numpy(), **d)
    # the stacked encoder (model/model.py:195-198) at a small size
    torch.manual_seed(11)
    enc = torch.nn.Sequential(BidirectionalLSTM(48, 32, 32), BidirectionalLSTM(32, 32, 32)).double()
    x = torch.randn(5, 6, 48, dtype=torch.float64, requires_grad=True)
    out = enc(x)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    d = {f"p.{k}": v.detach().numpy() for k, v in enc.state_dict().items()}
    d.update({f"g.{k}": v.grad.numpy() for k, v in enc.named_parameters()})
    np.savez_compressed(os.path.join(OUT, "encrnn_small.npz"), x=x.detach().numpy(),
                        out=out.detach().numpy(), w=w.numpy(), gx=x.grad.numpy(), **d)


def gen_ctc():
    g = torch.Generator().manual_seed(4321)

    def case(name, T, N, C, max_l, blank=0, var_in=False, peaky=False, lens=None, zero_len=False, big=False):
        x = torch.randn(T, N, C, generator=g, dtype=torch.float64) * (6.0 if peaky else 1.0)
        x.requires_grad_(True)
        tl = torch.randint(0 if zero_len else 1, max_l + 1, (N,), generator=g) if lens is None else torch.tensor(lens)
        labels = [c for c in range(C) if c != blank]
        tg = torch.zeros(N, max(int(tl.max()), 1), dtype=torch.long)
        for n in range(N):
            idx = torch.randint(0, len(labels), (int(tl[n]),), generator=g)
            tg[n, : int(tl[n])] = torch.tensor(labels)[idx] if int(tl[n]) else tg[n, :0]
        # force some repeated labels
        for n in range(0, N, 3):
            if int(tl[n]) >= 2:
                tg[n, 1] = tg[n, 0]
        il = torch.full((N,), T, dtype=torch.long)
        if var_in:
            il = torch.randint(1, T + 1, (N,), generator=g)
        out = dict(x=x.detach().numpy(), targets=tg.numpy(), input_lengths=il.numpy(),
                   target_lengths=tl.numpy(), blank=blank)
        lp = x.log_softmax(2)
        for zi in (False, True):
            nll = F.ctc_loss(lp, tg, il, tl, blank=blank, reduction="none", zero_infinity=zi)
            out[f"nll_zi{int(zi)}"] = nll.detach().numpy()
            for red in ("mean", "sum"):
                if big and not (red == "mean" and zi):
                    continue
                loss = F.ctc_loss(lp, tg, il, tl, blank=blank, reduction=red, zero_infinity=zi)
                gx, = torch.autograd.grad(loss, x, retain_graph=True)
                out[f"loss_{red}_zi{int(zi)}"] = loss.detach().numpy()
                out[f"grad_{red}_zi{int(zi)}"] = gx.numpy()
        # concatenated-target form must give the same numbers
        cat = torch.cat([tg[n, : int(tl[n])] for n in range(N)])
        loss_cat = F.ctc_loss(lp, cat, il, tl, blank=blank, reduction="mean", zero_infinity=True)
        assert torch.allclose(loss_cat, torch.as_tensor(out["loss_mean_zi1"])), name
        out["targets_concat"] = cat.numpy()
        np.savez_compressed(os.path.join(OUT, f"ctc_{name}.npz"), **out)

    case("small", 12, 6, 5, 4)
    case("varlen", 20, 9, 11, 6, var_in=True)            # includes infeasible samples (inf / NaN grads)
    case("peaky", 24, 5, 8, 8, peaky=True)
    case("blank2", 10, 4, 6, 3, blank=2)
    case("emptytgt", 8, 5, 4, 3, zero_len=True, lens=[0, 2, 0, 3, 1])
    case("tight", 6, 4, 5, 0, lens=[6, 3, 5, 4])          # T == L (+repeats -> infeasible rows)
    case("cfgb", 64, 3, 195, 32, big=True)                # cfg-B geometry, reduced batch


def gen_attention():
    """model/model.py:50-148  Attention (reference code, imported): greedy probs and teacher-forced logits
    in eval() mode, float32 CPU, plus the state dict the outputs were made with."""
    from model.model import Attention

    def case(name, B, T, C, H, V, steps, seed, blank=3, scale=1.0, store_weights=True):
        torch.manual_seed(seed)
        m = Attention(input_size=C, hidden_size=H, num_classes=V, sos_id=1, eos_id=2, pad_id=0, blank_id=blank,
                      dropout_p=0.1, sampling_prob=0.0).eval()
        with torch.no_grad():
            for p in m.parameters():          # larger weights: peaky attention and well separated argmax margins
                p.mul_(scale)
        g = torch.Generator().manual_seed(seed + 1)
        batch_H = torch.randn(B, T, C, generator=g)
        text = torch.randint(0, V, (B, steps), generator=g)
        text[:, 0] = 1
        with torch.no_grad():
            probs = m(batch_H, is_train=False, batch_max_length=steps - 1)
            logits = m(batch_H, text=text, is_train=True, batch_max_length=steps - 1)
        out = {"batch_H": batch_H.numpy(), "text": text.numpy(), "probs": probs.numpy(), "logits": logits.numpy(),
               "dims": np.array([B, T, C, H, V, steps, -1 if blank is None else blank])}
        if store_weights:
            for k, v in m.state_dict().items():
                out["sd." + k] = v.numpy()
        else:
            # large case: the weights are re-created from the seed by the build's own module, whose parameter
            # registration order (hence default init) equals the reference's -- checked here, once
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
            import rcnn_ocr_b200 as R
            torch.manual_seed(seed)
            ours = R.Attention(C, H, V, 1, 2, 0, blank, dropout_p=0.1)
            for k, v in m.state_dict().items():
                assert torch.equal(ours.state_dict()[k] * scale, v), k
            out["seed_scale"] = np.array([seed, scale], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, f"attn_{name}.npz"), **out)

    case("tiny", 3, 5, 64, 64, 20, 6, seed=11, scale=3.0)
    case("cfga", 16, 16, 256, 256, 194, 26, seed=12, scale=2.0, store_weights=False)   # configs/config.json geometry
    case("noblank", 5, 7, 64, 128, 30, 4, seed=13, blank=None, scale=3.0)


def gen_attention_train():
    """model/model.py:110-148 in train() mode (reference code, imported): teacher-forced logits and the gradients of every
    parameter and of batch_H for loss = sum(logits * w).  dropout_p = 0 (dropout draws cannot be reproduced across
    implementations); sampling_prob 0, 1 (every step feeds back its own argmax) and 0.5 (the reference draws
    torch.rand(1) on the CPU generator once per step: same seed, same decisions)."""
    from model.model import Attention

    def case(name, B, T, C, H, V, steps, seed, sampling, blank=3, scale=2.0):
        torch.manual_seed(seed)
        m = Attention(input_size=C, hidden_size=H, num_classes=V, sos_id=1, eos_id=2, pad_id=0, blank_id=blank,
                      dropout_p=0.0, sampling_prob=sampling).train()
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(scale)
        g = torch.Generator().manual_seed(seed + 1)
        batch_H = torch.randn(B, T, C, generator=g).requires_grad_(True)
        text = torch.randint(4, V, (B, steps), generator=g)
        text[:, 0] = 1
        w = torch.randn(B, steps, V, generator=g) / (B * steps) ** 0.5
        torch.manual_seed(seed + 2)               # the scheduled-sampling draws start here
        logits = m(batch_H, text=text, is_train=True, batch_max_length=steps - 1)
        (logits * w).sum().backward()
        out = {"batch_H": batch_H.detach().numpy(), "text": text.numpy(), "w": w.numpy(), "logits": logits.detach().numpy(),
               "g.batch_H": batch_H.grad.numpy(), "seed": seed, "sampling": sampling,
               "dims": np.array([B, T, C, H, V, steps, -1 if blank is None else blank])}
        # the weights are re-created from the seed by the build's own module (same registration order, hence the same
        # default init as the reference's: asserted here), times `scale`
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        import rcnn_ocr_b200 as R
        torch.manual_seed(seed)
        ours = R.Attention(C, H, V, 1, 2, 0, blank, dropout_p=0.0)
        for k, v in m.state_dict().items():
            assert torch.equal(ours.state_dict()[k] * scale, v), k
        out["scale"] = scale
        for k, p in m.named_parameters():
            out["g." + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"attn_train_{name}.npz"), **out)

    case("tf", 6, 9, 64, 64, 24, 7, seed=21, sampling=0.0)
    case("sampled", 5, 8, 32, 32, 24, 6, seed=22, sampling=1.0)
    case("mixed", 7, 6, 32, 64, 30, 8, seed=23, sampling=0.5, blank=None)


def _import_reference_transforms():
    """data/transforms.py imports albumentations (absent here) for the augmentation pipeline; ResizeAndPadA.apply and
    pack_attention_targets need only cv2 / numpy / torch, so the package is stubbed for the import."""
    import types
    if "albumentations" not in sys.modules:
        A = types.ModuleType("albumentations")

        class ImageOnlyTransform:
            def __init__(self, always_apply=True, p=1.0):
                pass

        A.ImageOnlyTransform = ImageOnlyTransform
        for name in ("Compose", "Normalize", "ShiftScaleRotate", "RandomBrightnessContrast", "InvertImg"):
            setattr(A, name, lambda *a, **k: None)
        pt = types.ModuleType("albumentations.pytorch")
        pt.ToTensorV2 = lambda *a, **k: None
        sys.modules["albumentations"] = A
        sys.modules["albumentations.pytorch"] = pt
    import data.transforms as T
    return T


def gen_preprocess():
    """Input step goldens from the reference's own ResizeAndPadA.apply (data/transforms.py:91-120) on seeded images of
    assorted sizes / channel counts, followed by A.Normalize(0.5, 0.5)'s arithmetic (albumentations' normalize:
    float32 (img - 127.5) * reciprocal(127.5)) and ToTensorV2's HWC -> CHW; and target packing goldens from
    pack_attention_targets (data/transforms.py:123-157)."""
    T = _import_reference_transforms()
    rng = np.random.default_rng(2024)
    for name, (ih, iw) in (("preproc_32x128", (32, 128)), ("preproc_64x256", (64, 256))):
        shapes = [(20, 57, 3), (11, 40, 3), (64, 200, 3), (50, 173, 3), (96, 300, 3), (32, 128, 3), (31, 90, 1), (70, 41, 4),
                  (ih, iw, 3), (2 * ih, 2 * iw, 3), (3 * ih, iw * 3, 3), (9, 900, 3), (200, 30, 3), (1, 1, 3), (ih // 2, iw // 2, 3)]
        if ih == 64:
            shapes = [(40, 150, 3), (130, 500, 3), (64, 256, 1), (20, 300, 4), (128, 512, 3)]
        tf = T.ResizeAndPadA(img_h=ih, img_w=iw)
        arrs, outs = {}, []
        for i, (h, w, c) in enumerate(shapes):
            # smooth + noise content (pure noise hides coordinate errors behind rounding)
            yy, xx = np.mgrid[0:h, 0:w]
            base = 127 + 90 * np.sin(xx / 5.0 + i) * np.cos(yy / 3.0) + rng.normal(0, 25, (h, w))
            img = np.clip(np.stack([base + 10 * k for k in range(c)], -1), 0, 255).astype(np.uint8)
            if c == 1:
                img = img[:, :, 0]
            arrs[f"img{i}"] = img
            canvas = tf.apply(img.copy())
            assert canvas.shape == (ih, iw, 3) and canvas.dtype == np.uint8
            outs.append(canvas)
        # the padded canvases as the reference produced them (uint8 HWC); Normalize + ToTensorV2 are applied by the test:
        # float32 (v - 127.5) * reciprocal(127.5), HWC -> CHW (albumentations' normalize, ToTensorV2's transpose)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), n=len(shapes), img_h=ih, img_w=iw, canvas=np.stack(outs), **arrs)
    itos = charset()
    stoi = {s: i for i, s in enumerate(itos)}
    texts = ["", "a", "привет мир", "Hello, World!", "x" * 40, "№42 «тест»", "\u4e2d\u6587 not in charset", " ", "a<b>c"]
    for ml in (25, 5):
        ti, ty, ln = T.pack_attention_targets(texts, stoi, ml)
        np.savez_compressed(os.path.join(OUT, f"pack_attn_ml{ml}.npz"), texts=np.array(texts, dtype=object), max_len=ml,
                            itos=np.array(itos, dtype=object),
                            text_in=ti.numpy(), target_y=ty.numpy(), lengths=ln.numpy())


if __name__ == "__main__":
    gen_decode()
    gen_decode_fn()
    gen_cnn()
    gen_bilstm()
    gen_ctc()
    gen_attention()
    gen_attention_train()
    gen_preprocess()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {tot / 1e6:.2f} MB -> {OUT}")
