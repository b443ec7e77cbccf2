#!/usr/bin/env python
"""Timing of the fused residual-GEMM + LayerNorm kernel over K (tuning aid): T(K) = rounds * (K/64 * t_kblock + t_epilogue)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

dev = torch.device("cuda:0")
M, N = 50432, 384
for K in (64, 128, 384, 768, 1536):
    A = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(3)]
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=dev), torch.ones(N, device=dev), torch.zeros(N, device=dev)
    X = [torch.randn(M, N, device=dev) for _ in range(3)]
    Xn = torch.empty(M, N, device=dev, dtype=torch.bfloat16)

    def fn(i):
        check(lib().smk_gemm_ln(ptr(A[i % 3]), K, ptr(W), ptr(bias), ptr(X[i % 3]), ptr(gamma), ptr(beta), ptr(Xn), M, N, K, 1e-6, stream_ptr()))
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"K={K:5d}: {us:7.1f} us   {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
