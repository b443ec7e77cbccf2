#!/usr/bin/env python
"""Phase timeline of CTA 0 of the attention kernel (clock64 stamps), to see where an item's time goes."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr
dev = torch.device("cuda:0")
B, N, H = 256, 197, 6
D = H * 64
qkv = (torch.randn(B * N, 3 * D, device=dev)).to(torch.bfloat16)
out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    check(lib().smk_attention_tc(ptr(qkv), ptr(out), B, N, H, 0.125, stream_ptr()))
tr = torch.zeros(16 * 12 * 8, dtype=torch.int64, device=dev)
check(lib().smk_debug_attn_trace(ptr(tr)))
check(lib().smk_attention_tc(ptr(qkv), ptr(out), B, N, H, 0.125, stream_ptr()))
torch.cuda.synchronize()
check(lib().smk_debug_attn_trace(None))
t = tr.cpu().view(16, 12, 8)
t0 = int(t[t > 0].min())
names = {0: "TMA ", 1: "MMA "}
for it in range(2, 6):
    print(f"--- item {it}")
    for w in [0, 1] + list(range(4, 12)):
        row = [int(x) - t0 if x > 0 else -1 for x in t[it, w]]
        role = names.get(w, f"S{(w - 4) // 4}q{w % 4}")
        print(f"  w{w} {role}: " + " ".join(f"{x:7d}" for x in row[:7]))
print("per-item period (MMA ev0):", [int(t[i + 1, 1, 0] - t[i, 1, 0]) for i in range(1, 9)])
