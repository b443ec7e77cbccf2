#!/usr/bin/env python
"""Device timing of the fp16s-mode encoder GEMM shapes: 3-term fp16 split (smk_gemm_split) against the fp8-corrected form
(smk_gemm_q8).  CUDA events, operands rotated through buffers larger than L2.  Tuning aid:
  SMK_GEMM_CTA_PAIR=0/1, SMK_GEMM_SWAP_AB=0/1, SMK_GEMM_BN=..., and with the tune build SMK_GEMM_DEBUG=1 (no epilogue) / 2 (no loads).
Usage: python scripts/gemm_q8_bench.py [--batch 256] [--only fc1,fc2]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
M = args.batch * 197


def timeit(fn, nbuf, iters=args.iters):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def case(name, N, K, epi, kind3, kind8, Mrows=None):
    Mr = Mrows or M
    nbuf = 4
    A = [torch.randn(Mr, 2 * K, device=dev).to(torch.float16) * 0.01 for _ in range(nbuf)]
    W = (torch.randn(N, 2 * K, device=dev) * 0.01).to(torch.float16)
    bias = torch.randn(N, device=dev)

    def outs(kind):
        return [torch.zeros(Mr, N * (1 if kind in (0, 1) else 2), device=dev, dtype=torch.float32 if kind == 1 else torch.float16) for _ in range(nbuf)]
    o3, o8 = outs(kind3), outs(kind8)
    ao, wo = (C.c_int32 * 3)(0, 0, K), (C.c_int32 * 3)(0, K, 0)

    def f3(i):
        check(lib().smk_gemm_split(ptr(A[i]), 2 * K, ptr(W), 2 * K, ptr(bias), ptr(o3[i]), o3[i].shape[1], Mr, N, K, epi, kind3, 1, 3, ao, wo, stream_ptr()), name)

    def f8(i):
        check(lib().smk_gemm_q8(ptr(A[i]), 2 * K, ptr(W), 2 * K, ptr(bias), ptr(o8[i]), o8[i].shape[1], Mr, N, K, epi, kind8, stream_ptr()), name)
    t3, t8 = timeit(f3, nbuf), timeit(f8, nbuf)
    gf = 2.0 * Mr * N * K / 1e6
    print(f"{name:6s} M={Mr} N={N:5d} K={K:5d}: 3-term fp16 {t3:7.1f} us ({gf / t3:6.1f} TF/s algorithmic)   fp8-corrected {t8:7.1f} us ({gf / t8:6.1f} TF/s)")


CASES = {"pe": (384, 768, 0, 1, 1, args.batch * 196), "proj": (384, 384, 4, 1, 1, None), "fc1": (1536, 384, 1, 3, 4, None),
         "fc1p": (1536, 384, 1, 0, 0, None), "fc2": (384, 1536, 4, 1, 1, None)}
def qkv_case():
    """qkv of the fp16s mode: A_hi·(W_hi + W_lo), fp16 output (not a q8 contraction)."""
    N, K, nbuf = 1152, 384, 4
    A = [torch.randn(M, K, device=dev).to(torch.float16) * 0.01 for _ in range(nbuf)]
    W = (torch.randn(N, 2 * K, device=dev) * 0.01).to(torch.float16)
    bias = torch.randn(N, device=dev)
    o = [torch.zeros(M, N, device=dev, dtype=torch.float16) for _ in range(nbuf)]
    ao, wo = (C.c_int32 * 3)(0, 0, 0), (C.c_int32 * 3)(0, K, 0)
    t = timeit(lambda i: check(lib().smk_gemm_split(ptr(A[i]), K, ptr(W), 2 * K, ptr(bias), ptr(o[i]), N, M, N, K, 0, 0, 1, 2, ao, wo, stream_ptr()), "qkv"), nbuf)
    print(f"qkv    M={M} N={N:5d} K={K:5d}: 2-term fp16 {t:7.1f} us ({2.0 * M * N * K / 1e6 / t:6.1f} TF/s algorithmic)")


def pe_tokens_case():
    """patch embed with token assembly (per-image tiles, position embedding, 3-D output map), fp8-corrected operands."""
    n_img, hw, N, K, nbuf = args.batch, 196, 384, 768, 4
    A = [torch.randn(n_img * hw, 2 * K, device=dev).to(torch.float16) * 0.01 for _ in range(nbuf)]
    W = (torch.randn(N, 2 * K, device=dev) * 0.01).to(torch.float16)
    bias, pos = torch.randn(N, device=dev), torch.randn(hw + 1, N, device=dev)
    o = [torch.zeros(n_img, hw + 1, N, device=dev) for _ in range(nbuf)]
    t = timeit(lambda i: check(lib().smk_gemm_tokens(ptr(A[i]), 2 * K, ptr(W), 2 * K, ptr(bias), ptr(pos), ptr(o[i]), N, n_img, hw, N, K, 2, stream_ptr()), "pe"), nbuf)
    print(f"pe_tok M={n_img * hw} N={N:5d} K={K:5d}: fp8-corrected, token assembly {t:7.1f} us ({2.0 * n_img * hw * N * K / 1e6 / t:6.1f} TF/s algorithmic)")


if not args.only or "qkv" in args.only.split(","):
    qkv_case()
if not args.only or "pe_tok" in args.only.split(","):
    pe_tokens_case()
for nm, (N, K, epi, k3, k8, mr) in CASES.items():
    if not args.only or nm in args.only.split(","):
        case(nm, N, K, epi, k3, k8, mr)
