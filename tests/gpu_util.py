"""Shared helpers for the GPU parity tests (all calls go through the C-ABI via selfmask_b200)."""
import ctypes as C

import numpy as np
import torch

import selfmask_b200 as S
from selfmask_b200._lib import check, lib, ptr, stream_ptr

DEV = torch.device("cuda:0")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def gemm_f32(A, W, bias, epi=0, C_init=None):
    M, K = A.shape
    N = W.shape[0]
    out = C_init.clone() if C_init is not None else torch.empty(M, N, dtype=torch.float32, device=DEV)
    check(lib().smk_gemm_f32(ptr(A), K, ptr(W), ptr(bias), ptr(out), N, M, N, K, epi, stream_ptr()), "gemm_f32")
    return out


def gemm_bf16(A, W, bias, epi=0, out_f32=False, C_init=None):
    M, K = A.shape
    N = W.shape[0]
    if C_init is not None:
        out = C_init.clone()
    else:
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=DEV)
    check(lib().smk_gemm_bf16(ptr(A), K, ptr(W), ptr(bias), ptr(out), N, M, N, K, epi, 1 if out_f32 else 0, stream_ptr()), "gemm_bf16")
    return out


def make_model(nq=20, mode="fp32", max_batch=4, return_intermediate=True, seed=0):
    from oracle import selfmask_oracle as O
    cfg = O.make_config(n_queries=nq)
    sd = O.synth_state_dict(cfg, seed=seed)
    m = S.SelfMaskB200(n_queries=nq, mode=mode, max_batch=max_batch, return_intermediate=return_intermediate).to(DEV)
    m.load_state_dict(sd)
    return m, sd, cfg


def forward_with_logits(model, x):
    """Run the model and also fetch the pre-sigmoid mask logits (debug tap)."""
    B, _, H, W = x.shape
    cfg = model.cfg
    L = cfg.dec_layers if model.return_intermediate else 1
    hp, wp = -(-H // cfg.patch), -(-W // cfg.patch)
    logits = torch.zeros(B, L, cfg.n_queries, hp * cfg.scale_factor, wp * cfg.scale_factor, dtype=torch.float32, device=DEV)
    handle = model._handle(B, H, W)
    check(lib().smk_model_debug_logits(handle, ptr(logits)), "debug_logits")
    out = model(x)
    check(lib().smk_model_debug_logits(handle, None), "debug_logits")
    torch.cuda.synchronize()
    return out, logits


def binarised_iou_agreement(p_a: np.ndarray, p_b: np.ndarray) -> np.ndarray:
    """IoU between the binarised (>0.5) masks of two implementations, per (image, query); 1.0 when both empty."""
    a, b = p_a > 0.5, p_b > 0.5
    inter = (a & b).sum(axis=(-1, -2)).astype(np.float64)
    union = (a | b).sum(axis=(-1, -2)).astype(np.float64)
    return np.where(union == 0, 1.0, inter / np.maximum(union, 1))
