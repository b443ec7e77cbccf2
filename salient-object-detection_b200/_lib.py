"""ctypes binding of the C-ABI in include/selfmask_b200.h.

The shared library is the product: if it is missing or a call fails, this module raises — there is no
CPU or PyTorch fallback anywhere in the package.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMK_LIB_PATH") or os.path.join(HERE, "libselfmask_b200.so")   # override: A/B runs of two builds (tuning)

SMK_MODE_FP32, SMK_MODE_BF16, SMK_MODE_BF16X3, SMK_MODE_FP16S = 0, 1, 2, 3
EPI_NONE, EPI_GELU, EPI_RELU, EPI_RESIDUAL = 0, 1, 2, 4
QCOUNT_STRIDE, MCOUNT_STRIDE, MSUM_STRIDE = 2, 528, 32


class SmkConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("patch", "dim", "depth", "heads", "mlp_dim", "n_queries", "dec_layers", "dec_ffn",
                                         "scale_factor", "pos_grid")]


class SmkError(RuntimeError):
    pass


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
# every symbol include/selfmask_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "smk_last_error": (C.c_char_p, []),
    "smk_version": (_I, []),
    "smk_launch_count": (_L, []),
    "smk_prof_enable": (_I, [_I]),
    "smk_prof_read": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_L)]),
    "smk_prof_timeline": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), _I]),
    "smk_prof_timeline2": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), _I]),
    "smk_model_debug_logits": (_I, [_P, _P]),
    "smk_weight_count": (_I, [C.POINTER(SmkConfig)]),
    "smk_weight_entry": (_I, [C.POINTER(SmkConfig), _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_L)]),
    "smk_weights_numel": (_L, [C.POINTER(SmkConfig)]),
    "smk_model_workspace_bytes": (_L, [C.POINTER(SmkConfig), _I, _I, _I, _I]),
    "smk_model_create": (_I, [C.POINTER(SmkConfig), _I, _P, _P, _L, _I, _I, _I, _P, C.POINTER(_P)]),
    "smk_model_destroy": (_I, [_P]),
    "smk_model_forward": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "smk_model_forward_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "smk_model_tap": (_I, [_P, _I, _P, _L, _P]),
    "smk_eval_batch": (_I, [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "smk_finalize_records": (_I, [_P, _P, _L, _F, _F, _F, _P, _P]),
    "smk_mask_metrics": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "smk_upsample_bilinear": (_I, [_P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "smk_gemm_f32": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "smk_gemm_bf16": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "smk_layernorm_f16": (_I, [_P, _P, _P, _P, _L, _P, _L, _I, _F, _I, _P]),
    "smk_im2col_f16": (_I, [_P, _I, _P, _I, _I, _I, _I, C.POINTER(C.c_float), _I, _P]),
    "smk_gemm_tokens": (_I, [_P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "smk_gemm_q8": (_I, [_P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "smk_split_q8": (_I, [_P, _L, _P, _L, _I, _I, _P]),
    "smk_gemm_split": (_I, [_P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "smk_gemm_batched": (_I, [_P, _L, _L, _I, _I, _I, _P, _L, _L, _I, _I, _I, _P, _I, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "smk_xattn_tc": (_I, [_P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "smk_xattn_fold_weights": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "smk_attention_tc_f16": (_I, [_P, _P, _L, _I, _I, _I, _I, _F, _P]),
    "smk_attention_small_f16": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _P, _L, _I, _I, _I, _I, _I, C.c_float, _I, _P]),
    "smk_dec_self_attention": (_I, [_P, _L, _P, _L, _P, _P, _I, _I, _I, _F, _P]),
    "smk_layernorm": (_I, [_P, _P, _P, _P, _L, _I, _F, _I, _P]),
    "smk_attention": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _L, _L, _F, _I, _P]),
    "smk_attention_tc": (_I, [_P, _P, _I, _I, _I, _F, _P]),
    "smk_attention_tc_multi": (_I, [_P, _L, _P, _L, _P, _L, _L, _L, _I, _I, _I, _P, _L, _I, _I, _I, _I, _I, _F, _I, _P]),
    "smk_attention_tc_general": (_I, [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "smk_attention_small": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _P, _L, _I, _I, _I, _I, _I, C.c_float, _P]),
    "smk_attention_fa": (_I, [_P, _P, _L, _P, _P, _L, _P, _P, _L, _I, _I, _I, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "smk_debug_attn_trace": (_I, [_P]),
    "smk_debug_gemm_trace": (_I, [_P]),
    "smk_debug_xattn_trace": (_I, [_P]),
    "smk_gemm_ln": (_I, [_P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "smk_split3": (_I, [_P, _L, _I, _P, _I, _P]),
    "smk_cast_bf16": (_I, [_P, _P, _L, _P]),
}

_lib = None


def lib():
    """Load libselfmask_b200.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmkError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "selfmask_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().smk_last_error().decode(errors="replace")
        raise SmkError(f"{what or 'selfmask_b200 call'} failed with status {status}: {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None → NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise SmkError(f"{name} must be a CUDA tensor (selfmask_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise SmkError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise SmkError(f"{name} must be contiguous")
    return t
