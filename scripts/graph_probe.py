#!/usr/bin/env python
"""Does replaying one forward + evaluation step from a CUDA graph beat stream launches?  (tuning probe)"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import selfmask_b200 as S  # noqa: E402
from selfmask_b200 import synthetic as Y  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = S.SelfMaskB200(n_queries=20, mode="bf16", max_batch=B).to(dev)
model.load_state_dict(Y.synth_state_dict(model.table(), seed=0))
uniq = min(B, 32)
x = torch.from_numpy(Y.synth_images_u8(uniq, 224, 224, seed=1234)).repeat((B + uniq - 1) // uniq, 1, 1, 1)[:B].to(dev)
g = torch.from_numpy(Y.synth_gt(uniq, 224, 224, seed=4321)).repeat((B + uniq - 1) // uniq, 1, 1, 1)[:B].to(dev)
rec = S.BatchRecords(B, 20, dev)


def step():
    out = model(x)
    S.eval_batch(out["mask_pred"], out["objectness"], g, up=4, out=rec)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"stream launches: {timeit(step):.3f} ms/step")
ref = rec.m_counts.clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
print(f"graph replay   : {timeit(graph.replay):.3f} ms/step")
torch.cuda.synchronize()
print("records identical:", bool(torch.equal(ref, rec.m_counts)))
