"""ctypes binding of include/rcnn_ocr_b200.h.  Fails loudly when the CUDA library is absent."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCNN_OCR_B200_LIB: load a differently built copy of the library (A/B experiments)
_LIB_PATH = os.path.abspath(os.environ.get("RCNN_OCR_B200_LIB") or os.path.join(_HERE, "librcnn_ocr_b200.so"))
_lib = None


class LibraryMissing(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def _declare(L: ctypes.CDLL) -> None:
    i, i64, vp, sz = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t
    L.rcnn_version.restype = i
    L.rcnn_last_error.restype = ctypes.c_char_p
    L.rcnn_device_check.restype = i
    L.rcnn_ctc_greedy.restype = i
    L.rcnn_ctc_greedy.argtypes = [vp, i, i, i, i, i64, i64, i, vp, vp, vp, vp]
    L.rcnn_ctc_workspace_bytes.restype = sz
    L.rcnn_ctc_workspace_bytes.argtypes = [i, i, i, i]
    L.rcnn_ctc_loss.restype = i
    L.rcnn_ctc_loss.argtypes = [vp, i, i, i, i, i64, i64, vp, i64, vp, vp, i, i, i, i,
                                vp, vp, vp, i64, i64, vp, sz, vp]
    L.rcnn_gemm_bf16.restype = i
    L.rcnn_gemm_bf16.argtypes = [vp, i64, vp, i64, vp, i64, i, vp, i, i, i, vp]
    L.rcnn_lstm_packed_bytes.restype = sz
    L.rcnn_lstm_packed_bytes.argtypes = [i, i]
    L.rcnn_lstm_pack_weights.restype = i
    L.rcnn_lstm_pack_weights.argtypes = [vp] * 8 + [i, i, vp, vp]
    L.rcnn_lstm_forward.restype = i
    L.rcnn_lstm_forward.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp]
    L.rcnn_lstm_backward.restype = i
    L.rcnn_lstm_backward.argtypes = [vp, vp, vp, vp, i, i, i, vp, vp, vp, ctypes.c_size_t, vp]
    L.rcnn_lstm_backward_workspace_bytes.restype = ctypes.c_size_t
    L.rcnn_lstm_backward_workspace_bytes.argtypes = [i, i, i]
    L.rcnn_colsum_bf16.restype = i
    L.rcnn_colsum_bf16.argtypes = [vp, i64, i64, i, vp, vp]
    L.rcnn_cast_bf16_2d.restype = i
    L.rcnn_cast_bf16_2d.argtypes = [vp, i64, vp, i64, i64, i, vp]
    L.rcnn_edit_distance.restype = i
    L.rcnn_edit_distance.argtypes = [vp, i64, vp, vp, vp, vp, i, vp, vp, i, i, vp, vp, vp, vp]
    L.rcnn_attn_score_context.restype = i
    L.rcnn_attn_score_context.argtypes = [vp, vp, vp, vp, i64, i64, i, i, i, i, vp, vp, i64, vp]
    L.rcnn_attn_cell.restype = i
    L.rcnn_attn_cell.argtypes = [vp, vp, vp, i, i, i, vp, vp, i64, i, vp, i64, vp]
    L.rcnn_attn_argmax.restype = i
    L.rcnn_attn_argmax.argtypes = [vp, i, i, i, vp, i64, vp, vp]
    L.rcnn_attn_score_context_ld.restype = i
    L.rcnn_attn_score_context_ld.argtypes = [vp, vp, i64, vp, vp, i64, i64, i, i, i, i, vp, vp, i64, vp]
    L.rcnn_attn_score_context_bf16.restype = i
    L.rcnn_attn_score_context_bf16.argtypes = [vp, vp, i64, vp, vp, i64, i64, i, i, i, i, vp, vp, i64, vp]
    L.rcnn_attn_step_bf16.restype = i
    L.rcnn_attn_step_bf16.argtypes = [vp, vp, i64, vp, vp, i64, i64, i, i, i, i, vp, vp, i64, vp, i64, i, i, vp, i64, vp, vp]
    L.rcnn_attn_gates_cell.restype = i
    L.rcnn_attn_gates_cell.argtypes = [vp, i64, vp, i64, vp, vp, vp, i, i, i, i, vp, vp, i64, vp, i64, vp]
    L.rcnn_chain_launches.restype = i
    L.rcnn_chain_launches.argtypes = [i]
    L.rcnn_attn_step_train.restype = i
    L.rcnn_attn_step_train.argtypes = [vp, vp, i64, vp, vp, i64, i64, i, i, i, i, vp, vp, vp, i64, vp]
    L.rcnn_attn_gates_cell_train.restype = i
    L.rcnn_attn_gates_cell_train.argtypes = [vp, i64, vp, i64, vp, vp, vp, i, i, i, i, vp, vp, vp, i64, vp, i64, vp, vp]
    L.rcnn_attn_cell_bwd.restype = i
    L.rcnn_attn_cell_bwd.argtypes = [vp, vp, vp, vp, i64, vp, i64, vp, i, i, vp, i64, vp]
    L.rcnn_attn_step_bwd.restype = i
    L.rcnn_attn_step_bwd.argtypes = [vp, i64, vp, vp, vp, i64, i64, vp, vp, i64, vp, i, i, i, i, vp, vp, i64, vp, vp]
    L.rcnn_attn_dprojH.restype = i
    L.rcnn_attn_dprojH.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp]
    L.rcnn_attn_greedy_decode.restype = i
    L.rcnn_attn_greedy_decode.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, i, vp]
    L.rcnn_attn_train_forward.restype = i
    L.rcnn_attn_train_forward.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.rcnn_attn_train_backward.restype = i
    L.rcnn_attn_train_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.rcnn_se_gate_workspace_bytes.restype = ctypes.c_size_t
    L.rcnn_se_gate_workspace_bytes.argtypes = [i, i]
    L.rcnn_se_gate.restype = i
    L.rcnn_se_gate.argtypes = [vp, i, i, i, i, vp, vp, i, vp, vp, vp, vp]
    L.rcnn_se_apply.restype = i
    L.rcnn_se_apply.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp, vp]
    L.rcnn_maxpool2x2_nhwc.restype = i
    L.rcnn_maxpool2x2_nhwc.argtypes = [vp, i, i, i, i, i, vp, vp]
    L.rcnn_attn_argmax_ld.restype = i
    L.rcnn_attn_argmax_ld.argtypes = [vp, i64, i, i, i, vp, i64, vp, vp]
    L.rcnn_lstm_forward_fused.restype = i
    L.rcnn_lstm_forward_fused.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    L.rcnn_preprocess_lines.restype = i
    L.rcnn_preprocess_lines.argtypes = [vp, vp, i, i, i, i, i, vp, i, vp]
    L.rcnn_reserve_sms.restype = i
    L.rcnn_reserve_sms.argtypes = [i]
    L.rcnn_lstm_plan.restype = i
    L.rcnn_lstm_plan.argtypes = [i, i, i, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.rcnn_lstm_weight_grads.restype = i
    L.rcnn_lstm_weight_grads.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, i, vp]
    L.rcnn_lstm_pack_weights_parts.restype = i
    L.rcnn_lstm_pack_weights_parts.argtypes = [vp] * 8 + [i, i, vp, i, vp]
    L.rcnn_launch_count.restype = ctypes.c_ulonglong
    L.rcnn_debug_timeline.restype = i
    L.rcnn_debug_timeline.argtypes = [vp]
    L.rcnn_debug_refetch_counter.restype = i
    L.rcnn_debug_refetch_counter.argtypes = [vp]
    L.rcnn_lstm_hprev.restype = i
    L.rcnn_lstm_hprev.argtypes = [vp, vp, i, i, i, vp]
    L.rcnn_lstm_unpack_grads.restype = i
    L.rcnn_lstm_unpack_grads.argtypes = [vp, vp, vp, i, i] + [vp] * 8 + [vp]
    L.rcnn_cast_bf16_3d.restype = i
    L.rcnn_cast_bf16_3d.argtypes = [vp, i64, i64, i64, vp, i, i, i, vp]
    L.rcnn_transpose_bf16.restype = i
    L.rcnn_transpose_bf16.argtypes = [vp, i64, vp, i64, i, i, vp]
    L.rcnn_gemm_bf16_atb.restype = i
    L.rcnn_gemm_bf16_atb.argtypes = [vp, i64, vp, i64, vp, i64, i, i, i, i, vp]
    L.rcnn_gemm_bf16_atb_grouped.restype = i
    L.rcnn_gemm_bf16_atb_grouped.argtypes = [vp, i64, i, vp, i64, i, vp, i64, i64, i, i, i, i, i, vp]
    L.rcnn_prof_enable.restype = i
    L.rcnn_prof_enable.argtypes = [i]
    L.rcnn_prof_reset.restype = i
    L.rcnn_prof_read.restype = i
    L.rcnn_prof_read.argtypes = [i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    L.rcnn_ctc_scale_grad.restype = i
    L.rcnn_ctc_scale_grad.argtypes = [vp, i, i, i, i64, i64, vp, i, vp]


def lib() -> ctypes.CDLL:
    """The loaded shared library.  Raises LibraryMissing if it has not been built --
    there is deliberately no other execution path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LibraryMissing(
                f"{_LIB_PATH} not found: build it with `python rcnn-ocr_b200/build.py` "
                "(or __graft_entry__.build()); rcnn-ocr_b200 has no CPU or eager fallback")
        L = ctypes.CDLL(_LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rcnn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (got {t.device}); rcnn-ocr_b200 has no CPU path")


def prof_enable(on: bool) -> None:
    check(lib().rcnn_prof_enable(int(on)), "rcnn_prof_enable")
    lib().rcnn_prof_reset()


def prof_read(kernel: int):
    """(total_ms, launches) of a dominant kernel since the last reset (CUDA events on the
    launching stream; see include/rcnn_ocr_b200.h)."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int(0)
    check(lib().rcnn_prof_read(int(kernel), ctypes.byref(ms), ctypes.byref(n)), "rcnn_prof_read")
    return ms.value, n.value
