"""ctypes binding of libtgfr_b200.so (C ABI in include/tgfr_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 the import /
first call raises.  Build with ``python -m text_guided_face_recognition_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libtgfr_b200.so")

PREC_FP32 = 0
PREC_TC = 1

P = c_void_p     # every tensor argument is passed as a raw device address
I = c_int
L = c_int64
F = c_float
Z = c_size_t

# name -> (restype, argtypes); mirrors include/tgfr_b200.h one to one
SIGNATURES = {
    "tgfr_version": (I, []),
    "tgfr_last_error": (c_char_p, []),
    "tgfr_device_check": (I, []),
    "tgfr_wordregion_fwd": (I, [P, L, L, L, P, L, L, L, P, I, I, I, I, I, F, F, F, F, P, P, I, I, P, Z, P, Z, P]),
    "tgfr_wordregion_bwd": (I, [P, L, L, L, P, L, L, L, P, I, I, I, I, I, F, F, F, F, P, P, P, I, P, Z, P, Z, P]),
    "tgfr_wordregion_workspace_bytes": (Z, [I, I, I, I, I, I]),
    "tgfr_wordregion_saved_bytes": (Z, [I, I, I, I, I, I]),
    "tgfr_attention_fwd": (I, [P, L, L, L, P, L, L, L, I, I, I, I, F, P, P, P]),
    "tgfr_attention_bwd": (I, [P, L, L, L, P, L, L, L, I, I, I, I, F, P, P, P, P, P]),
    "tgfr_cosine_scores_fwd": (I, [P, L, P, L, I, I, I, F, I, F, P, P, I, P, P, P, P]),
    "tgfr_cosine_scores_bwd": (I, [P, L, P, L, I, I, I, F, I, F, P, P, P, P, P, P, Z, P]),
    "tgfr_cosine_workspace_bytes": (Z, [I, I, I]),
    "tgfr_pair_ce_stats": (I, [P, I, I, I, P, P, P, P, P]),
    "tgfr_pair_ce_finish": (I, [P, P, P, P, I, I, I, F, P, P, P]),
    "tgfr_pair_ce_bwd": (I, [P, P, P, P, P, I, I, I, F, P, P]),
    "tgfr_cos_logits_fwd": (I, [P, L, P, L, L, I, I, I, F, I, P, L, P, P, I, P, Z, P]),
    "tgfr_arc_margin_apply": (I, [P, L, P, I, I, I, F, F, I, P, P]),
    "tgfr_arc_margin_bwd": (I, [P, L, P, L, L, P, P, P, P, P, L, I, I, I, I, F, F, I, P, P, I, P, Z, P]),
    "tgfr_margin_workspace_bytes": (Z, [I, I, I, I]),
    "tgfr_arc_fused_workspace_bytes": (Z, [I, I, I]),
    "tgfr_arc_fused_saved_bytes": (Z, [I, I, I]),
    "tgfr_arc_fused_fwd": (I, [P, L, P, L, L, P, I, I, I, I, F, F, I, P, P, P, P, P, P, P, Z, P, Z, P]),
    "tgfr_arc_fused_bwd": (I, [P, L, P, L, L, P, P, P, P, P, P, I, I, I, I, F, F, I, P, P, P, Z, P, Z, P]),
    "tgfr_mag_margin_fwd": (I, [P, P, I, I, F, I, P, P]),
    "tgfr_mag_margin_bwd": (I, [P, P, P, P, I, I, F, I, P, P, P]),
    "tgfr_cos_logits_bwd": (I, [P, L, P, L, L, P, P, P, L, P, L, I, I, I, F, I, P, P, I, P, Z, P]),
    "tgfr_ce_rows_stats": (I, [P, L, P, I, I, I, P, P, P, P]),
    "tgfr_focal_finish": (I, [P, P, P, I, F, P, P, P]),
    "tgfr_ce_rows_bwd": (I, [P, L, P, P, P, P, I, I, I, P, L, P]),
    "tgfr_cosine_rows_fwd": (I, [P, L, L, P, L, L, L, I, F, P, P, P]),
    "tgfr_cosine_rows_bwd": (I, [P, L, L, P, L, L, L, I, F, P, P, P, P, P]),
    "tgfr_merge_softmax_stats": (I, [P, I, I, I, P, P]),
    "tgfr_mag_ce_stats": (I, [P, P, L, P, I, I, P, P, P, P, P]),
    "tgfr_mag_ce_bwd": (I, [P, P, L, P, P, P, I, I, P, P, P]),
    "tgfr_texthead_saved_bytes": (Z, [I, I, I, I]),
    "tgfr_texthead_workspace_bytes": (Z, [I, I, I, I]),
    "tgfr_texthead_fwd": (I, [P, P, P, P, P, P, P, I, I, I, I, I, P, P, P, Z, P]),
    "tgfr_texthead_bwd": (I, [P, P, P, I, I, I, I, I, P, P, P, P, P, P, P, Z, P, Z, P]),
    "tgfr_pair_cosine": (I, [P, L, L, P, L, L, L, I, F, P, P]),
    "tgfr_roc_workspace_bytes": (Z, [L]),
    "tgfr_roc_curve": (I, [P, P, L, I, P, P, P, P, P, Z, P]),
    "tgfr_row_argmax": (I, [P, L, I, I, P, P]),
    "tgfr_fcfm_working_num_params": (I, []),
    "tgfr_fcfm_working_fwd": (I, [P, L, L, L, L, P, L, L, L, P, L, P, L, P, I, I, I, P, L, P]),
    "tgfr_fcfm_working_workspace_bytes": (Z, [I]),
    "tgfr_fcfm_working_fwd_tc": (I, [P, L, L, L, L, P, L, L, L, P, L, P, L, P, I, I, I, P, L, P, Z, P]),
    "tgfr_imim_saved_bytes": (Z, [I, I]),
    "tgfr_imim_workspace_bytes": (Z, [I, I]),
    "tgfr_imim_num_params": (I, []),
    "tgfr_imim_fwd": (I, [P, L, L, L, P, I, I, I, I, F, F, P, P, P, P, Z, P]),
    "tgfr_imim_bwd": (I, [P, P, P, L, L, L, P, I, I, I, I, P, Z, P, P, P, Z, P]),
    "tgfr_matmul_split_workspace_bytes": (Z, [I, I, I, I, I]),
    "tgfr_matmul_split": (I, [I, P, L, P, L, P, L, I, I, I, I, F, P, I, I, I, P, Z, P]),
    "tgfr_proj_head_fwd": (I, [P, L, P, P, I, I, I, P, P, P]),
    "tgfr_proj_head_bwd": (I, [P, P, P, P, L, P, I, I, I, P, P, P, P, P]),
    "tgfr_fcfm_train_saved_bytes": (Z, [I, I]),
    "tgfr_fcfm_train_workspace_bytes": (Z, [I, I]),
    "tgfr_fcfm_train_fwd": (I, [P, L, L, L, L, P, L, L, P, L, P, L, P, I, I, I, I, F, F, P, P, L, P, Z, P]),
    "tgfr_fcfm_train_bwd": (I, [P, L, P, L, L, P, L, P, L, P, I, I, I, I, P, Z, P, P, P, P, P, P, Z, P]),
    "tgfr_debug_umma": (I, [P, P, P, I, I, I, I, I, P]),
    "tgfr_debug_tma_reduce": (I, [P, I, I, P]),
    "tgfr_debug_umma_2cta": (I, [P, P, P, I, I, P]),
    "tgfr_debug_set_trace": (I, [P]),
}

_lib = None
_checked_devices = set()


class TgfrError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TgfrError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -m text_guided_face_recognition_b200.build` (needs nvcc with sm_100a support). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tgfr_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise TgfrError(f"{what} failed (code {rc}): {last_error()}")


def ensure_device(device):
    """Raise unless `device` is a CUDA device this library supports (sm_100)."""
    import torch
    if device.type != "cuda":
        raise TgfrError(
            f"tensor on {device}: text_guided_face_recognition_b200 runs on CUDA (sm_100a) only; "
            "there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream (stream_ptr); a tensor that lives elsewhere would be
        # read through a foreign context.  nn.DataParallel's replica threads and torch.cuda.device(...) set this right.
        raise TgfrError(f"tensor on cuda:{idx} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                        "wrap the call in `with torch.cuda.device(tensor.device)`")
    if idx in _checked_devices:
        return
    with torch.cuda.device(idx):
        check(load().tgfr_device_check(), "tgfr_device_check")
    _checked_devices.add(idx)


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()
