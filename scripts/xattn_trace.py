#!/usr/bin/env python
"""Phase trace (clock64) of CTA 0 of the restructured cross-attention kernel (tuning aid).  Usage: python scripts/xattn_trace.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import selfmask_b200 as S  # noqa: E402
from selfmask_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
nq, H, D, hw = 20, 6, 384, 196
N = hw + 1
qp = (torch.randn(B * nq, H * D, device=dev) * 0.25).half()
tok = (torch.randn(B * N, D, device=dev) * 1.2).half()
out = torch.empty(B * nq, 2 * H * D, device=dev, dtype=torch.float16)
buf = torch.zeros(8 * 2 * 8, dtype=torch.int64, device=dev)
for _ in range(3):
    check(lib().smk_xattn_tc(ptr(qp), ptr(tok), N, 1, ptr(out), B, nq, H, D, hw, stream_ptr()))
torch.cuda.synchronize()
lib().smk_debug_xattn_trace(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
check(lib().smk_xattn_tc(ptr(qp), ptr(tok), N, 1, ptr(out), B, nq, H, D, hw, stream_ptr()))
e1.record()
torch.cuda.synchronize()
lib().smk_debug_xattn_trace(None)
t = buf.cpu().view(8, 2, 8)
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us for {B} images")
t0 = int(t[0, 0, 0])
names_m = ["loop top", "tmem free", "T landed", "S issued", "P ready", "U issued"]
names_s = ["loop top", "S ready", "P written", "U ready", "stored"]
for it in range(2):
    if int(t[it, 0, 0]) == 0:
        break
    print(f"image {it}: MMA warp   " + "  ".join(f"{n} {int(t[it, 0, i]) - t0}" for i, n in enumerate(names_m)))
    print(f"image {it}: softmax w2 " + "  ".join(f"{n} {int(t[it, 1, i]) - t0}" for i, n in enumerate(names_s)))
