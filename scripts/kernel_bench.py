#!/usr/bin/env python
"""Per-kernel device timing of the encoder-shaped launches (CUDA events, L2 flushed between iterations by rotating
through buffers larger than L2).  Usage: python scripts/kernel_bench.py [--batch 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--shapes", default="", help="extra GEMM cases name:M:N:K:epi:f32,...")
args = ap.parse_args()
dev = torch.device("cuda:0")
M = args.batch * 197
PEAK = 1374.4


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
    return e0.elapsed_time(e1) / iters * 1e3   # us


def gemm_case(name, N, K, epi, f32, M=None):
    M = M or globals()["M"]
    nbuf = 4
    A = [(torch.randn(M, K, device=dev)).to(torch.bfloat16) for _ in range(nbuf)]
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    Cs = [torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(nbuf)]

    def fn(i):
        check(lib().smk_gemm_bf16(ptr(A[i]), K, ptr(W), ptr(bias), ptr(Cs[i]), N, M, N, K, epi, 1 if f32 else 0, stream_ptr()), name)
    us = timeit(fn, nbuf)
    tf = 2.0 * M * N * K / us / 1e6
    byts = M * K * 2 + M * N * (4 if f32 else 2) * (2 if epi & 4 else 1)
    print(f"{name:8s} M={M} N={N:5d} K={K:5d} epi={epi} {'f32' if f32 else 'bf16'}: {us:8.1f} us  {tf:7.1f} TFLOP/s ({tf / PEAK:.2f} of sustained)  "
          f"min-HBM {byts / 1e6:6.0f} MB = {byts / us / 1e3:6.0f} GB/s")


CASES = {"qkv": (1152, 384, 0, False), "proj": (384, 384, 4, True), "fc1": (1536, 384, 1, False), "fc2": (384, 1536, 4, True),
         "kv": (4608, 384, 0, False)}
for name, c in CASES.items():
    if not args.only or name in args.only.split(","):
        gemm_case(name, *c)
for sh in filter(None, args.shapes.split(",")):
    nm, m_, n_, k_, e_, f_ = sh.split(":")
    gemm_case(nm, int(n_), int(k_), int(e_), bool(int(f_)), M=int(m_))
if args.only or args.shapes:
    sys.exit(0)

# attention (encoder shape)
B, N, H = args.batch, 197, 6
D = H * 64
qkv = [(torch.randn(B * N, 3 * D, device=dev)).to(torch.bfloat16) for _ in range(3)]
out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
us = timeit(lambda i: check(lib().smk_attention_tc(ptr(qkv[i]), ptr(out), B, N, H, 0.125, stream_ptr())), 3)
fl = 4.0 * N * N * 64 * H * B
print(f"attn     B={B} N={N}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s ({fl / us / 1e6 / PEAK:.2f})")

# layernorm
x = [torch.randn(M, D, device=dev) for _ in range(3)]
g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
us = timeit(lambda i: check(lib().smk_layernorm(ptr(x[i]), ptr(g), ptr(b), ptr(y), M, D, 1e-6, 1, stream_ptr())), 3)
print(f"layernorm rows={M}: {us:8.1f} us  {M * D * 6 / us / 1e3:7.0f} GB/s")
