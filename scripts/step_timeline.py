#!/usr/bin/env python
"""Per-launch timeline of one forward+evaluation step from the library's CUDA-event profiler (tuning aid).
Events between launches break PDL overlap, so the sum is a little above the un-profiled step; shares are what matter.
Usage: python scripts/step_timeline.py [--batch 256] [--mode bf16] [--size 224]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import selfmask_b200 as S  # noqa: E402
from selfmask_b200._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--mode", default="bf16")
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--nq", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
from selfmask_b200 import synthetic as Y  # noqa: E402  (same synthetic weights / images / ground truth as bench.py)
model = S.SelfMaskB200(n_queries=args.nq, mode=args.mode, max_batch=args.batch).to(dev)
model.load_state_dict(Y.synth_state_dict(model.table(), seed=0))
uniq = min(args.batch, 32)
rep = (args.batch + uniq - 1) // uniq
x = torch.from_numpy(Y.synth_images_u8(uniq, args.size, args.size, seed=1234)).repeat(rep, 1, 1, 1)[:args.batch].to(dev)
gt = torch.from_numpy(Y.synth_gt(uniq, args.size, args.size, seed=4321)).repeat(rep, 1, 1, 1)[:args.batch].to(dev)
for _ in range(3):
    out = model(x)
    rec = S.eval_batch(out["mask_pred"], out["objectness"], gt)
torch.cuda.synchronize()
lib().smk_prof_enable(1)
out = model(x)
rec = S.eval_batch(out["mask_pred"], out["objectness"], gt)
torch.cuda.synchronize()
cap = 4096
ms, cat, st = (C.c_float * cap)(), (C.c_int * cap)(), (C.c_float * cap)()
n = lib().smk_prof_timeline(ms, cat, st, cap)
lib().smk_prof_enable(0)
names = ["gemm_tc", "attn_simt", "gemm_f32", "layernorm", "eval", "mask_head", "other", "attn_tc"]
tot = 0.0
for i in range(n):
    tot += ms[i]
    print(f"{i:4d} {st[i] * 1e3:9.1f} us  {names[cat[i]]:10s} {ms[i] * 1e3:8.1f} us")
print(f"launches {n}  sum {tot:.3f} ms  span {st[n - 1] + ms[n - 1]:.3f} ms")
