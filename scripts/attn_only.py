#!/usr/bin/env python
"""Encoder-shaped attention launches only (ncu target)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr
dev = torch.device("cuda:0")
B, N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 197, 6
D = H * 64
qkv = (torch.randn(B * N, 3 * D, device=dev)).to(torch.bfloat16)
out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    check(lib().smk_attention_tc(ptr(qkv), ptr(out), B, N, H, 0.125, stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
