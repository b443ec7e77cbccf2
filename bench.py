#!/usr/bin/env python
"""Benchmark of the SelfMask inference + evaluation hot path (BASELINE.json metric: images/s at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference --gpus N ...            # the reference algorithm's CPU path (oracle port)

A step = one pass of the hot path over one batch of synthetic images per GPU: model forward (encoder, decoder,
mask head, objectness; all 6 decoder layers' masks, as the reference computes them) + the fused evaluation
(x4 upsample, per-query IoU, query selection, IoU / F-measure / MAE / pixel-acc / S-measure reductions) and, at
N > 1, the one collective on the path (all-reduce of the per-image count rows).  Workload = BASELINE.json
configs[2]: nq 20, 224x224, batch 256 per GPU (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--nq", type=int, default=20)
    ap.add_argument("--mode", default="fp16s", choices=["fp16s", "bf16", "bf16x3", "fp32"],
                    help="numeric mode of the headline; fp16s (default) is the tensor-core mode inside north_star's 2e-2 / 99.9 %% tolerance")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-throughput-mode", action="store_true", help="skip the extra single-pass bf16 figure (outside the tolerance)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the live fp32-validation-mode comparison of the benchmarked mode")
    ap.add_argument("--sweep", type=int, default=0,
                    help="BASELINE.json configs[4]: evaluate a DUTS-TE-shaped sweep of this many synthetic images (5019) through "
                         "Evaluator.__call__, sharded over the ranks, and print the sweep line instead of the step benchmark")
    ap.add_argument("--cpu-sample", type=int, default=256, help="images in the bounded CPU-baseline sample")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: warm up, then run exactly ONE step between cudaProfilerStart/Stop and exit "
                         "(ncu --profile-from-start off sees the step's launches 0..n-1 in order; prints no bench line)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML in-process
    (the library nvidia-smi itself reads; ~1 ms per sample, so even a 50 ms timed region gets dozens of samples); falls
    back to an `nvidia-smi -lms` child process when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake_slowdown": 0x80}

    def __init__(self, index):
        self.index, self.proc, self.lines, self.nvml, self.samples = index, None, [], None, []
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)      # first calls are slow (tens of ms): prime them
            try:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                except Exception:
                    pw = float("nan")
                self.samples.append((float(mhz), int(mask), pw))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
            sm = [x[0] for x in self.samples]
            pw = [x[2] for x in self.samples if x[2] == x[2]]
            reasons = sorted(k for k, bit in self.BITS.items() if any(x[1] & bit for x in self.samples))
            try:
                limit_w = self.nvml.nvmlDeviceGetEnforcedPowerLimit(self.handle) / 1e3
            except Exception:
                limit_w = None
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(sm), "source": "nvml",
                    "power_w_nvml_avg": float(np.median(pw)) if pw else None, "power_limit_w": limit_w,
                    "power_note": "NVML reports a ~1 s running average, so a 50 ms timed region after an idle phase reads low"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [s.strip() for s in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_reference_images_per_s(n_img, size, nq, batch, steps=1, warmup=0):
    """The reference algorithm on the host cores: oracle port (torch CPU fp32 model + the reference's metric
    semantics, per-image loop like evaluator.pyc).  Returns (images/s, cores, description)."""
    from oracle import selfmask_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.make_config(n_queries=nq)
    sd = O.synth_state_dict(cfg, seed=0)
    xs = O.normalize_images(O.synth_images_u8(n_img, size, size, seed=1234))
    gts = O.synth_gt(n_img, size, size, seed=4321)
    batches = [(xs[i:i + batch], gts[i:i + batch]) for i in range(0, n_img, batch)]

    def one_pass():
        with torch.no_grad():
            O.evaluate(lambda x: O.model_forward(sd, x, cfg), batches)
    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return n_img * steps / dt, cores, dt / steps


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (the reference is Python/PyTorch and
    cannot travel to the GPU box, so this is the oracle port; kind = "port").  Rank 0 only."""
    if rank != 0:
        return
    sample = min(args.cpu_sample, 16)
    ips, cores, sec = cpu_reference_images_per_s(sample, args.size, args.nq, batch=8, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": f"SelfMask nq{args.nq} {args.size}x{args.size} images/sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, sample_images_per_step=sample),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} images/step (batches of 8): oracle model forward + reference metric loop, torch CPU fp32, "
                                       f"{cores} threads"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, **extra):
    c = {"workload": f"SelfMask nq{args.nq} ViT-S/16 inference+metrics, batch {args.batch}/GPU at {args.size}x{args.size} "
                     f"(BASELINE.json configs[2]), batch-sharded, one record gather per step",
         "per_gpu_batch": args.batch, "image": [args.size, args.size], "n_queries": args.nq, "numeric_mode": args.mode,
         "mask_layers": 6, "input": "uint8 pixels (normalisation fused on device)",
         "images": f"{args.batch} distinct synthetic images per rank (no tiling)",
         "l2": "working set larger than L2: every step streams > 1 GB of activations through the 126 MB L2, so the 51 MB of "
               "inputs (x uint8 + GT) are evicted between steps"}
    if args.mode == "fp16s":
        c["numeric_mode_detail"] = ("fp16 tcgen05 operands, fp32 accumulate (TMEM), fp32 residual stream / LayerNorm / softmax; patch embed, proj, fc1, "
                                    "fc2 = fp16 hi·hi + the two correction products hi·lo + lo·hi on e4m3 operands (fp8 tensor-core rate, scale-input-d); "
                                    "qkv = A_hi·(W_hi + W_lo); encoder attention single-pass fp16; decoder self-attention fp32 on the CUDA cores; "
                                    "cross-attention restructured fp16 on tcgen05; decoder tail / heads = 3-term bf16 splits")
    c.update(extra)
    return c


TAG_NAMES = ["other", "patch_embed", "qkv", "attention", "proj", "fc1", "fc2", "layernorm", "memory_kv", "decoder_gemm", "decoder_attention",
             "decoder_layernorm", "mask_logits", "mask_upsample", "objectness", "eval_query_iou", "eval_mask_metrics", "im2col"]
TENSOR_TAGS = {1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 14}  # rows whose yardstick is the tensor pipe (12: the mask-logit contraction, a batched tcgen05 GEMM since round 2); the rest are HBM rows


def per_kernel_rows(lib, C, step, n_prof, pk, ms_plain):
    """Per-kernel roofline rows from the library's CUDA-event profiler (one event pair per launch on the launching stream).
    achieved = ALGORITHMIC work (SURVEY.md §8d FLOPs / bytes, credited once whatever number of split terms the tensor core is
    issued) ÷ the kernel's summed launch time; `issued_frac` is the same with the FLOPs really issued.  Event pairs between
    launches defeat the programmatic-dependent-launch overlap, so the rows sum to more than the un-profiled step: both are printed."""
    lib.smk_prof_enable(1)
    for _ in range(n_prof):
        step(sync=True)
    torch.cuda.synchronize()
    cap = 16384
    ms, cat, tag = (C.c_float * cap)(), (C.c_int * cap)(), (C.c_int * cap)()
    work, issued = (C.c_double * cap)(), (C.c_double * cap)()
    n = lib.smk_prof_timeline2(ms, cat, tag, work, issued, cap)
    lib.smk_prof_enable(0)
    rows = {}
    for i in range(max(n, 0)):
        t = tag[i] if 0 <= tag[i] < len(TAG_NAMES) else 0
        if t == 0 and cat[i] == 4:                       # evaluation kernels are launched outside the model: split them by order
            t = 15 if rows.get("_eval_toggle", 0) % 2 == 0 else 16
            rows["_eval_toggle"] = rows.get("_eval_toggle", 0) + 1
        if t == 0 and cat[i] == 5:
            t = 12 if rows.get("_mask_toggle", 0) % 2 == 0 else 13
            rows["_mask_toggle"] = rows.get("_mask_toggle", 0) + 1
        r = rows.setdefault(t, {"ms": 0.0, "work": 0.0, "issued": 0.0, "launches": 0})
        r["ms"] += ms[i]; r["work"] += work[i]; r["issued"] += issued[i]; r["launches"] += 1
    rows.pop("_eval_toggle", None); rows.pop("_mask_toggle", None)
    out, total = {}, sum(r["ms"] for r in rows.values())
    for t, r in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
        tensor = t in TENSOR_TAGS
        sec = r["ms"] / 1e3
        ach = r["work"] / sec / (1e12 if tensor else 1e9) if sec > 0 else 0.0
        iss = r["issued"] / sec / (1e12 if tensor else 1e9) if sec > 0 else 0.0
        peak = pk["tensor"] if tensor else pk["hbm"]
        out[TAG_NAMES[t]] = {"ms_per_step": r["ms"] / n_prof, "launches_per_step": r["launches"] // n_prof,
                             "us_per_launch": 1e3 * r["ms"] / max(r["launches"], 1), "share": r["ms"] / total if total else 0.0,
                             "bound": "tensor" if tensor else "hbm", "unit": "TFLOP/s" if tensor else "GB/s",
                             "achieved": ach, "frac": ach / peak, "issued": iss, "issued_frac": iss / peak,
                             "work_per_launch": r["work"] / max(r["launches"], 1)}
    return out, total / n_prof


def parity_vs_fp32_mode(S, Y, model, dev, nq, size, n_img=8):
    """Live parity figure of a tensor-core mode: its mask logits / binarised masks / objectness top-1 against this library's own
    fp32 validation mode (CUDA-core fp32, pinned to the reference at <= 1e-4 by tests/test_gpu_model.py) on the same images.
    north_star: logits max-abs 2e-2, IoU agreement >= 99.9 %."""
    from selfmask_b200._lib import check, lib, ptr
    ref = S.SelfMaskB200(n_queries=nq, mode="fp32", max_batch=n_img, return_intermediate=True).to(dev)
    ref.load_state_dict(model.state_dict())
    x = torch.from_numpy(Y.synth_images_u8(n_img, size, size, seed=99)).to(dev)

    def run(m):
        cfg = m.cfg
        hp = -(-size // cfg.patch)
        lg = torch.zeros(n_img, cfg.dec_layers, cfg.n_queries, hp * cfg.scale_factor, hp * cfg.scale_factor, dtype=torch.float32, device=dev)
        h = m._handle(n_img, size, size)
        check(lib().smk_model_debug_logits(h, ptr(lg)), "debug_logits")
        out = m(x)
        check(lib().smk_model_debug_logits(h, None), "debug_logits")
        torch.cuda.synchronize()
        return out, lg
    o_a, l_a = run(model)
    o_b, l_b = run(ref)
    a, b = o_a["mask_pred"][:, -1] > 0.5, o_b["mask_pred"][:, -1] > 0.5
    inter, union = (a & b).sum((-1, -2)).double(), (a | b).sum((-1, -2)).double()
    agree = torch.where(union == 0, torch.ones_like(union), inter / union.clamp(min=1))
    top = (o_a["objectness"][:, -1, :, 0].argmax(-1) == o_b["objectness"][:, -1, :, 0].argmax(-1)).sum().item()
    res = {"against": "this library's fp32 validation mode (itself <= 1e-4 from the reference: tests/test_gpu_model.py)", "images": n_img,
           "logits_max_abs": float((l_a - l_b).abs().max()), "iou_agreement_mean": float(agree.mean()), "iou_agreement_min": float(agree.min()),
           "objectness_top1_match": f"{int(top)}/{n_img}", "north_star": {"logits_max_abs": 2e-2, "iou_agreement": 0.999}}
    res["meets_north_star"] = bool(res["logits_max_abs"] <= 2e-2 and res["iou_agreement_mean"] >= 0.999 and top == n_img)
    del ref
    return res


def sweep_batches(Y, n_total, start, stop, batch, size, pin=True):
    """Pinned uint8 batches of images [start, stop) of the synthetic DUTS-TE-shaped sweep: every image its own seed; an empty GT
    every 97th image, a full GT every 194th and one GT whose centroid lies on row 0 (the reference's NaN S-measure)."""
    out = []
    for b0 in range(start, stop, batch):
        idx = list(range(b0, min(b0 + batch, stop)))
        x = torch.empty(len(idx), 3, size, size, dtype=torch.uint8)
        m = torch.empty(len(idx), 1, size, size, dtype=torch.uint8)
        for j, i in enumerate(idx):
            x[j] = torch.from_numpy(Y.synth_images_u8(1, size, size, seed=100000 + i)[0])
            g = Y.synth_gt(1, size, size, seed=200000 + i)[0]
            if i % 97 == 96:
                g[:] = 0
            if i % 194 == 193:
                g[:] = 1
            if i == n_total // 2:
                g[:] = 0
                g[0, 0, 3:9] = 1          # all foreground on row 0 → centroid Y = 0 → empty quadrants → NaN S-measure in the reference
            m[j] = torch.from_numpy(g)
        out.append({"x": x.pin_memory() if pin else x, "m": m.pin_memory() if pin else m})
    return out


def run_sweep(args, S, Y, dev, rank, world, dist):
    """BASELINE.json configs[4]: the whole DUTS-TE-shaped sweep through Evaluator.__call__ (sharded by shard_range, one gather of
    the per-image records, ordered means on every rank).  The line carries a digest of the gathered records and the 14 averages,
    so that runs at different GPU counts can be compared bit for bit."""
    import hashlib
    n_total, B = args.sweep, args.batch
    a, b = S.shard_range(n_total, rank, world)
    model = S.SelfMaskB200(n_queries=args.nq, mode=args.mode, max_batch=B, return_intermediate=True).to(dev)
    model.load_state_dict(Y.synth_state_dict(model.table(), seed=0))
    batches = sweep_batches(Y, n_total, a, b, B, args.size)
    ev = S.Evaluator(network=model, dataset=batches)
    ev(dataset_name="duts_te_synthetic", dir_ckpt=None, batch_size=B, device=dev)      # warm-up sweep (allocator, handles)
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ev(dataset_name="duts_te_synthetic", dir_ckpt=None, batch_size=B, device=dev)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt.item()))
    recs = ev.global_records if world > 1 else ev.device_records()
    counts, sums = recs["m_counts"].cpu().numpy(), recs["m_sums"].cpu().numpy()
    digest = hashlib.sha256(counts.tobytes() + sums.tobytes()).hexdigest()
    ties = torch.tensor([ev.objectness_ties], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ties)
    if rank == 0:
        t = float(np.median(times))
        line = {"metric": f"SelfMask nq{args.nq} {args.size}x{args.size} images/sec", "kind": "sweep", "value": n_total / t, "unit": "images/s",
                "n_gpus": world, "sweep_images": n_total, "sweep_seconds_median": t, "sweep_seconds": times, "higher_is_better": True,
                "dtype": "f32" if args.mode == "fp32" else ("f16" if args.mode == "fp16s" else "bf16"), "data": "synthetic",
                "config": {"workload": f"DUTS-TE-shaped eval sweep: {n_total} synthetic images at {args.size}x{args.size}, nq{args.nq}, "
                                       f"{world} GPU(s), full IoU/F-measure/MAE/S-measure reduction (BASELINE.json configs[4])",
                           "numeric_mode": args.mode, "per_gpu_batch": B, "shards": [S.shard_range(n_total, r, world) for r in range(world)],
                           "edge_cases": "empty GT every 97th image, full GT every 194th, one NaN-S-measure GT (centroid on row 0)",
                           "timing": "wall clock around Evaluator.__call__ (pinned host batches, H2D inside), median of 3 sweeps, max over ranks"},
                "records_rows": int(counts.shape[0]), "records_sha256": digest, "objectness_top1_ties": int(ties.item()),
                "result": {k: (None if v != v else v) for k, v in res.items()}, "result_nan_keys": [k for k, v in res.items() if v != v]}
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    import selfmask_b200 as S
    from selfmask_b200 import synthetic as Y
    from selfmask_b200._lib import lib
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU reference arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.sweep > 0:
        run_sweep(args, S, Y, dev, rank, world, dist)
        if world > 1:
            dist.destroy_process_group()
        return
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    model = S.SelfMaskB200(n_queries=args.nq, mode=args.mode, max_batch=B, return_intermediate=True).to(dev)
    model.load_state_dict(Y.synth_state_dict(model.table(), seed=0))
    # B distinct images per rank (the evaluation kernels' work is data-dependent: no tiling of a small set).  Raw uint8 pixels: the
    # loader's ImageNet normalisation (datasets/base_dataset.py:250) is fused into the patch im2col on the device, bit-identically
    # (tests/test_gpu_model.py), so a step moves 1 byte per pixel-channel over PCIe instead of 4
    x_host = torch.from_numpy(Y.synth_images_u8(B, args.size, args.size, seed=1234 + rank)).contiguous().pin_memory()
    g_host = torch.from_numpy(Y.synth_gt(B, args.size, args.size, seed=4321 + rank)).contiguous().pin_memory()
    x, g = x_host.to(dev), g_host.to(dev)
    n_total = B * world
    rec = S.BatchRecords(B, args.nq, dev)
    exchange = S.RecordExchange(B, dev) if world > 1 else None
    pending = []

    def step(sync=False):
        """One pass of the hot path over one batch.  N > 1: the step's record gather is posted asynchronously (NCCL stream) and
        collected one step later, so it overlaps the next step's encoder; `sync` collects immediately."""
        out = model(x)
        S.eval_batch(out["mask_pred"], out["objectness"], g, up=4, out=rec)
        if exchange is None:
            return rec.m_counts, rec.m_sums
        pending.append(exchange.post(rec.m_counts, rec.m_sums))
        res = None
        while len(pending) > (0 if sync else 1):
            res = exchange.collect(pending.pop(0))
        return res

    def drain():
        res = None
        while pending:
            res = exchange.collect(pending.pop(0))
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step()
    drain()
    barrier()
    if args.ncu_step:
        torch.cuda.cudart().cudaProfilerStart()
        step(sync=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    sampler = ClockSampler(local_rank)
    launches0 = lib().smk_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    sampler.start()
    evs[0].record()
    last = None
    for i in range(K):
        r = step()
        last = r if r is not None else last
        evs[i + 1].record()
    if exchange is not None:
        last = drain() or last              # the final gather completes inside the timed region
        evs[K].record()
    barrier()
    ms = evs[0].elapsed_time(evs[K])
    step_ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(K))
    launches = lib().smk_launch_count() - launches0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_total * K / (ms / 1e3)
    counts, sums = last
    res = S.summarize(counts, sums)        # metric of the last step (sanity; outside the timed region)

    # ---- end to end through the public API: pinned host batches → Evaluator.__call__ → 14 averages ----------
    ev = S.Evaluator(network=model)

    def e2e_pass():
        ev.dataset = ({"x": x_host, "m": g_host} for _ in range(K))
        return ev(dataset_name="synthetic", dir_ckpt=None, batch_size=B, device=dev)    # N > 1: gathers the ranks' records itself
    e2e_pass()                                   # warm-up (allocator, page-locked paths)
    e2e_dts = []
    for _ in range(5):
        barrier()
        t0 = time.perf_counter()
        e2e_res = e2e_pass()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_dts.append(float(dt.item()))
    e2e_value = n_total * K / float(np.median(e2e_dts))
    h2d = x_host.numel() * x_host.element_size() + g_host.numel() * g_host.element_size()
    d2h = B * 2 * 8 * 8     # the finalised metric values (8 doubles per evaluated mask); the integer records stay on the device

    # ---- per-kernel device time with CUDA events on the launching stream (roofline block) -------------------------
    pk = peaks()
    rows, prof_ms = per_kernel_rows(lib(), C, step, 2, pk, ms / K)
    drain()
    dom = max(rows, key=lambda k: rows[k]["ms_per_step"])
    traffic = None          # real DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (args.batch, args.size, args.nq) == (256, 224, 20):
        traffic = json.load(open(tpath)).get(f"dram_bytes_per_launch_{args.mode}", {}).get(dom)
    d = rows[dom]
    # whole-step figure on the algorithmic FLOPs of SURVEY.md §8d (10.441 GFLOP per image at 224^2, nq 20)
    gflop_img = {(224, 20): 10.441, (224, 10): 10.165, (384, 20): 33.644}.get((args.size, args.nq))
    roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": pk["tensor"] if d["bound"] == "tensor" else pk["hbm"],
                "unit": d["unit"], "frac": d["frac"], "issued": d["issued"], "issued_frac": d["issued_frac"], "traffic": traffic,
                "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" if traffic is not None else None,
                "peak_source": f"MEASURED_PEAKS.json ({pk['src']}; sustained bf16 for a kernel timed inside a long step; burst {pk['tensor_burst']})",
                "avg_launch_ms": d["us_per_launch"] / 1e3,
                "definition": "achieved = algorithmic FLOPs (2·M·N·K of the contraction, SURVEY.md §8d; split-operand terms are NOT credited) "
                              "per launch / the kernel's CUDA-event launch time; issued = the same with the tensor-core terms really issued "
                              "(the e4m3 correction products of the fp16s GEMMs count as 2·M·N·2K issued at the fp8 rate, i.e. twice the peak this fraction is taken against)",
                "kernels": rows, "event_timed_ms_per_step": prof_ms, "plain_ms_per_step": ms / K,
                "note": "event pairs around every launch defeat the PDL overlap, so the per-kernel rows sum to more than the plain step",
                "limiter": ("the split GEMMs (fc1 / fc2 / proj / patch embed) are shared-memory-bandwidth bound, not MMA bound: with K = 384 every operand "
                            "byte of a 128 x 256 tile is written once by TMA and read once by the tensor core, 4 bytes per split element "
                            "(profiles/r02_gemm_q8.md: MMA only 79.5 us, + operand stream 101.7 us, + epilogue 133 us for fc1)") if args.mode == "fp16s" else None,
                "whole_step": ({"algorithmic_gflop_per_image": gflop_img, "achieved_tflops": value / world * gflop_img / 1e3,
                                "frac_of_sustained_bf16_peak": value / world * gflop_img / 1e3 / pk["tensor"]} if gflop_img else None)}

    if rank == 0:
        line = {"metric": f"SelfMask nq{args.nq} {args.size}x{args.size} images/sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": ms / K, "ms_per_step_median": step_ms[K // 2], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else ("f16" if args.mode == "fp16s" else "bf16"),
                "data": "synthetic", "config": workload_config(args),
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "passes_s": e2e_dts, "note": "median of 5 passes of K steps through Evaluator.__call__ (wall clock, max over ranks)"},
                "gpu_launches": int(launches), "roofline": roofline,
                "sanity": {"iou": res["iou"], "f_max": res["f_max"], "e2e_iou": e2e_res["iou"]}}
        if world == 1 and not args.no_parity_check and args.mode != "fp32":
            try:
                line["parity"] = parity_vs_fp32_mode(S, Y, model, dev, args.nq, args.size)
            except Exception as e:
                line["parity"] = {"error": str(e)[:200]}
        if world == 1 and args.mode == "fp16s" and not args.no_throughput_mode:
            # the same step in single-pass bf16 (every contraction one bf16 tcgen05 pass): the fastest mode, OUTSIDE north_star's
            # tolerance on these weights — reported with its own measured parity, never as the headline
            try:
                del ev
                mb = S.SelfMaskB200(n_queries=args.nq, mode="bf16", max_batch=B, return_intermediate=True).to(dev)
                mb.load_state_dict(Y.synth_state_dict(mb.table(), seed=0))

                def step_b():
                    ob = mb(x)
                    S.eval_batch(ob["mask_pred"], ob["objectness"], g, up=4, out=rec)
                for _ in range(3):
                    step_b()
                torch.cuda.synchronize()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(20):
                    step_b()
                p1.record()
                torch.cuda.synchronize()
                msb = p0.elapsed_time(p1) / 20
                line["throughput_mode"] = {"numeric_mode": "bf16", "value": B / (msb / 1e3), "unit": "images/s", "ms_per_step": msb, "steps": 20,
                                           "note": "single-pass bf16 operands, device-resident inputs; outside the 2e-2 logit tolerance (see parity)",
                                           "parity": parity_vs_fp32_mode(S, Y, mb, dev, args.nq, args.size)}
                del mb
            except Exception as e:      # never lose the headline line over the extra figure
                line["throughput_mode"] = {"numeric_mode": "bf16", "error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            ips, cores, sec = cpu_reference_images_per_s(args.cpu_sample, args.size, args.nq, batch=8)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} images (batches of 8) of the same workload: oracle model forward + the reference's "
                                              f"per-image metric loop, torch CPU fp32, {cores} threads, {sec:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
