#!/usr/bin/env python
"""Where the tcgen05 GEMM's roles wait (clock64 accounting, mean over CTAs): producer / MMA issuer / epilogue warp 2."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from selfmask_b200._lib import check, lib, ptr, stream_ptr
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="qkv:50432:1152:384:0:0,proj:50432:384:384:4:1,fc1:50432:1536:384:1:0,fc2:50432:384:1536:4:1")
ap.add_argument("--q8", action="store_true", help="fp16s-mode form: smk_gemm_q8 on [hi | q8] rows; the last shape field is then the output kind (0 / 1 / 3 / 4)")
args = ap.parse_args()
dev = torch.device("cuda:0")
for sh in args.shapes.split(","):
    nm, M, N, K, epi, f32 = sh.split(":")
    M, N, K, epi, f32 = int(M), int(N), int(K), int(epi), int(f32)
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    C = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    run = lambda: check(lib().smk_gemm_bf16(ptr(A), K, ptr(W), ptr(bias), ptr(C), N, M, N, K, epi, f32, stream_ptr()))
    if args.q8:
        A = (torch.randn(M, 2 * K, device=dev) * 0.01).to(torch.float16)
        W = (torch.randn(N, 2 * K, device=dev) * 0.01).to(torch.float16)
        C = torch.zeros(M, N * (1 if f32 in (0, 1) else 2), device=dev, dtype=torch.float32 if f32 == 1 else torch.float16)
        run = lambda: check(lib().smk_gemm_q8(ptr(A), 2 * K, ptr(W), 2 * K, ptr(bias), ptr(C), C.shape[1], M, N, K, epi, f32, stream_ptr()))
    for _ in range(3):
        run()
    tr = torch.zeros(16 * 148, dtype=torch.int64, device=dev)
    check(lib().smk_debug_gemm_trace(ptr(tr)))
    run()
    torch.cuda.synchronize()
    check(lib().smk_debug_gemm_trace(None))
    t = tr.cpu().view(148, 16).double()
    t = t[t[:, 8] > 0]
    m = torch.stack([t[:, i][t[:, i] > 0].mean() if (t[:, i] > 0).any() else torch.tensor(0.0, dtype=torch.float64) for i in range(16)])
    tiles = m[8]
    print(f"{nm} M={M} N={N} K={K}: {len(t)} CTAs, {tiles:.1f} tiles/CTA; per tile (cycles): "
          f"producer total {m[1]/tiles:.0f} (slot wait {m[0]/tiles:.0f}) | MMA total {m[4]/tiles:.0f} (operand wait {m[2]/tiles:.0f}, "
          f"accumulator wait {m[3]/tiles:.0f}) | epilogue total {m[7]/tiles:.0f} (accumulator wait {m[5]/tiles:.0f}, staging wait {m[6]/tiles:.0f}, tmem-ld wait {m[9]/tiles:.0f}, store section {m[10]/tiles:.0f})")
