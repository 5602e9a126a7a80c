"""CPU oracle for the RCNN-OCR sequence-recognition hot path.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  May be imported by tests/,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` -- never by the product package ``rcnn-ocr_b200/``.

The arithmetic lives in ``oracle.c`` (plain C, float64), loaded here with ctypes and
wrapped with numpy.  ``lstm_ref.py`` adds a torch-float64 explicit-equation BiLSTM whose
autograd gives the backward oracle, ``host_ref.py`` restates the small pure-Python
pieces of the reference (charset, token decode, CER/accuracy) and ``ref_port.py`` is
the timed CPU baseline (the reference's own torch-CPU op sequence).

Parity pinning: tests/golden/*.npz were generated in the build container by
tests/make_golden.py from the reference's own code (model/model.py,
training/utils.py imported from /root/reference) and from torch float64
``F.ctc_loss``; tests/test_oracle_golden.py checks this oracle against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (recipe: oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c"))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_int, c_ll, vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
        L.oracle_version.restype = c_int
        L.oracle_ctc_greedy_f32.restype = c_int
        L.oracle_ctc_greedy_f32.argtypes = [vp, c_int, c_int, c_int, c_ll, c_ll, c_int, vp, vp]
        for name in ("oracle_ctc_loss_f64", "oracle_ctc_loss_f32"):
            f = getattr(L, name)
            f.restype = c_int
            f.argtypes = [vp, c_int, c_int, c_int, c_int, c_ll, c_ll, vp, c_ll, vp, vp,
                          c_int, c_int, vp, vp]
        L.oracle_bilstm_forward_f64.restype = c_int
        L.oracle_bilstm_forward_f64.argtypes = [vp, c_int, c_int, c_int, c_int, c_int] + [vp] * 12
        _lib = L
    return _lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- decode

def ctc_greedy_ids(logits_btc: np.ndarray, blank: int = 0):
    """argmax / collapse repeats / strip blank over [B,T,C] float32 (any strides with a
    unit class stride).  Returns (ids [B,T] int32 padded with -1, lens [B] int32).
    Follows training/utils.py:135-150."""
    x = np.asarray(logits_btc, dtype=np.float32)
    assert x.ndim == 3
    if x.strides[2] != 4:
        x = np.ascontiguousarray(x)
    B, T, C = x.shape
    ids = np.full((B, T), -1, dtype=np.int32)
    lens = np.zeros((B,), dtype=np.int32)
    if B and T:
        rc = lib().oracle_ctc_greedy_f32(_ptr(x), B, T, C, x.strides[0] // 4, x.strides[1] // 4,
                                         int(blank), _ptr(ids), _ptr(lens))
        if rc:
            raise RuntimeError(f"oracle_ctc_greedy_f32 rc={rc}")
    return ids, lens


def ctc_greedy_decoder(logits, alphabet, blank: int = 0):
    """Same surface as the reference's training/utils.py:122-150, including its
    ``shape[0] < shape[1]`` layout heuristic (utils.py:132-133).  Returns (texts, seqs)."""
    x = np.asarray(logits, dtype=np.float32)
    if x.ndim == 3 and x.shape[0] < x.shape[1]:
        x = x.transpose(1, 0, 2)
    ids, lens = ctc_greedy_ids(x, blank)
    seqs = [ids[b, : lens[b]].tolist() for b in range(ids.shape[0])]
    texts = ["".join(alphabet[p - 1] for p in s) for s in seqs]
    return texts, seqs


# --------------------------------------------------------------------------- CTC

def ctc_nll_and_grad(x_tnc: np.ndarray, targets, input_lengths, target_lengths, blank: int = 0,
                     zero_infinity: bool = False, from_logits: bool = True, want_grad: bool = True,
                     threads: int = 1, f32: bool = False):
    """Per-sample nll [N] (float64) and per-sample gradient [T,N,C] (unscaled; see
    oracle.c for the convention).  ``targets`` is padded [N,Smax] or 1-D concatenated."""
    dt = np.float32 if f32 else np.float64
    x = np.ascontiguousarray(np.asarray(x_tnc, dtype=dt))
    T, N, C = x.shape
    tg = np.ascontiguousarray(np.asarray(targets, dtype=np.int64))
    il = np.ascontiguousarray(np.asarray(input_lengths, dtype=np.int64))
    tl = np.ascontiguousarray(np.asarray(target_lengths, dtype=np.int64))
    tstride = tg.shape[1] if tg.ndim == 2 else 0
    if tg.ndim == 2 and tstride == 0:
        tg = np.zeros((N, 1), dtype=np.int64)
        tstride = 1
    nll = np.zeros((N,), dtype=np.float64)
    grad = np.zeros((T, N, C), dtype=dt) if want_grad else None
    fn = lib().oracle_ctc_loss_f32 if f32 else lib().oracle_ctc_loss_f64
    es = x.itemsize

    def run(lo, hi):
        if hi <= lo:
            return 0
        # slice samples [lo,hi): x / grad are addressed through strides, targets by offset
        xs = x[:, lo:hi]
        if tstride:
            tgs = tg[lo:hi]
        else:
            off = int(tl[:lo].sum())
            tgs = tg[off: off + int(tl[lo:hi].sum())]
        tgs = np.ascontiguousarray(tgs)
        if tgs.size == 0:
            tgs = np.zeros((1,), dtype=np.int64)
        g = np.zeros((T, hi - lo, C), dtype=dt) if want_grad else None
        rc = fn(ctypes.c_void_p(xs.ctypes.data), int(from_logits), T, hi - lo, C,
                x.strides[0] // es, x.strides[1] // es, _ptr(tgs), tstride,
                _ptr(np.ascontiguousarray(il[lo:hi])), _ptr(np.ascontiguousarray(tl[lo:hi])),
                int(blank), int(zero_infinity),
                ctypes.c_void_p(nll[lo:hi].ctypes.data), _ptr(g))
        if want_grad:
            grad[:, lo:hi] = g
        return rc

    if threads <= 1 or N < 2 * threads:
        rcs = [run(0, N)]
    else:
        cuts = np.linspace(0, N, threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            rcs = list(ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])))
    if any(rcs):
        raise RuntimeError(f"oracle_ctc_loss rc={rcs}")
    return nll, grad


def ctc_loss(x_tnc, targets, input_lengths, target_lengths, blank: int = 0,
             reduction: str = "mean", zero_infinity: bool = False, from_logits: bool = True,
             want_grad: bool = True):
    """nn.CTCLoss semantics on top of ctc_nll_and_grad: returns (loss, grad) with the
    gradient of the REDUCED loss.  'mean' divides each nll by clamp(target_len, 1) and
    averages over the batch, 'sum' adds, 'none' returns [N] (grad is then of loss.sum())."""
    nll, grad = ctc_nll_and_grad(x_tnc, targets, input_lengths, target_lengths, blank,
                                 zero_infinity, from_logits, want_grad)
    tl = np.asarray(target_lengths, dtype=np.float64)
    N = nll.shape[0]
    if reduction == "mean":
        w = 1.0 / (np.maximum(tl, 1.0) * max(N, 1))
        loss = float((nll * w).sum()) if N else float("nan")
    elif reduction == "sum":
        w = np.ones_like(tl)
        loss = float(nll.sum())
    elif reduction == "none":
        w = np.ones_like(tl)
        loss = nll
    else:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    if grad is not None:
        grad = grad * w[None, :, None]
    return loss, grad


# --------------------------------------------------------------------------- BiLSTM

_LSTM_KEYS = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
              "rnn.weight_ih_l0_reverse", "rnn.weight_hh_l0_reverse",
              "rnn.bias_ih_l0_reverse", "rnn.bias_hh_l0_reverse", "linear.weight", "linear.bias")


def bilstm_forward(x_bti: np.ndarray, params: dict, return_hcat: bool = False):
    """BidirectionalLSTM.forward (model/model.py:159-163) in float64.  ``params`` uses
    the reference's state-dict keys (rnn.weight_ih_l0 ... linear.bias)."""
    x = np.ascontiguousarray(np.asarray(x_bti, dtype=np.float64))
    B, T, I = x.shape
    p = [np.ascontiguousarray(np.asarray(params[k], dtype=np.float64)) for k in _LSTM_KEYS]
    H = p[1].shape[1]
    O = p[8].shape[0]
    assert p[0].shape == (4 * H, I) and p[8].shape == (O, 2 * H)
    out = np.zeros((B, T, O), dtype=np.float64)
    hcat = np.zeros((B, T, 2 * H), dtype=np.float64)
    rc = lib().oracle_bilstm_forward_f64(_ptr(x), B, T, I, H, O, *[_ptr(a) for a in p],
                                         _ptr(out), _ptr(hcat))
    if rc:
        raise RuntimeError(f"oracle_bilstm_forward_f64 rc={rc}")
    return (out, hcat) if return_hcat else out
