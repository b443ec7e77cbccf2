"""Batched GPU evaluator with the reference `Evaluator` call surface.

Reference: `evaluator.pyc` (source not shipped): `Evaluator(network, arch, dir_dataset,
visualizer, debug)` and `__call__(dataset_name, dir_ckpt, img_size, scale_factor, batch_size, device,
cost_type) -> dict` of 14 averages + `{dir_ckpt}/metrics_{dataset_name}.txt`.

What changes is *where* the work happens: instead of ≥16 device→host syncs per image, one
`smk_eval_batch` call per batch does the x4 upsample, the per-query IoU counts, the query selection and
all metric reductions on the GPU; the host only turns integer counts into ratios and keeps the running
means in dataset order (bit-compatible with `AverageMeter`).
"""
import os
from typing import Callable, Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .metrics import METRIC_KEYS, finalize, finalize_device, running_mean, values_from_device


class BatchRecords:
    """Per-image integer / float records of an evaluated batch (device tensors)."""

    def __init__(self, B: int, nq: int, device):
        self.q_counts = torch.empty(B, nq, _lib.QCOUNT_STRIDE, dtype=torch.int32, device=device)
        self.idx = torch.empty(B, 2, dtype=torch.int32, device=device)
        self.m_counts = torch.empty(B, 2, _lib.MCOUNT_STRIDE, dtype=torch.int32, device=device)
        self.m_sums = torch.empty(B, 2, _lib.MSUM_STRIDE, dtype=torch.float64, device=device)


def eval_batch(mask_pred: torch.Tensor, objectness: torch.Tensor, gt: torch.Tensor, up: int = 4,
               out: Optional[BatchRecords] = None) -> BatchRecords:
    """evaluator.pyc@L199-226 for a whole batch.

    mask_pred: b x (L x) nq x h' x w' probabilities (5-D → last layer, @L199-205); objectness: b x (L x) nq (x 1);
    gt: b x 1 x H x W (or b x H x W) {0,1}, any integer/bool dtype.  Returns device-side records.
    """
    _lib.require_cuda(mask_pred, "mask_pred", torch.float32)
    if mask_pred.ndim == 5:
        mask_pred, objectness = mask_pred[:, -1], objectness[:, -1]
    objectness = objectness.reshape(objectness.shape[0], -1).float()
    B, nq, hp, wp = mask_pred.shape
    if mask_pred.stride()[1:] != (hp * wp, wp, 1) or objectness.stride(1) != 1:
        mask_pred, objectness = mask_pred.contiguous(), objectness.contiguous()
    g = gt.reshape(B, gt.shape[-2], gt.shape[-1])
    if g.dtype != torch.uint8:
        g = (g != 0).to(torch.uint8)
    g = g.to(mask_pred.device).contiguous()
    H, W = g.shape[-2:]
    rec = out or BatchRecords(B, nq, mask_pred.device)
    with torch.cuda.device(mask_pred.device):
        check(lib().smk_eval_batch(ptr(mask_pred), mask_pred.stride(0), ptr(objectness), objectness.stride(0), ptr(g),
                                   B, nq, hp, wp, up, H, W, ptr(rec.q_counts), ptr(rec.idx), ptr(rec.m_counts), ptr(rec.m_sums),
                                   stream_ptr()), "smk_eval_batch")
    return rec


def summarize(m_counts, m_sums) -> Dict[str, float]:
    """[n_img,2,528] / [n_img,2,32] records in dataset order → the reference's 14-key result
    (evaluator.pyc@L294-308).  Device tensors are finalised on the GPU (smk_finalize_records), numpy arrays on the
    host (metrics.finalize); both give identical bits."""
    if isinstance(m_counts, torch.Tensor) and m_counts.is_cuda:
        vals = values_from_device(finalize_device(m_counts, m_sums).cpu().numpy())
    else:
        vals = finalize(np.asarray(m_counts), np.asarray(m_sums))
    res = {k: running_mean(vals[k][:, 0]) for k in METRIC_KEYS}
    res.update({k + "_ub": running_mean(vals[k][:, 1]) for k in METRIC_KEYS})
    return res


def objectness_top1_ties(objectness: torch.Tensor) -> torch.Tensor:
    """Per image: does the largest objectness value occur more than once?  (device bool tensor, no synchronisation)

    The reference picks the mask with `torch.argsort(objectness, descending=True)[0]` (evaluator.pyc@L219-221), an UNSTABLE sort:
    on an exact tie of the top value the index it returns is data-dependent (an all-equal 20-vector gives 10),
    whereas this path takes the lowest index.  Ties cannot be reproduced by design, so they are detected and reported instead:
    `Evaluator.objectness_ties` counts them per sweep and parity harnesses exclude / flag those images."""
    o = objectness[:, -1] if objectness.ndim == 4 else objectness
    o = o.reshape(o.shape[0], -1)
    return (o == o.max(dim=1, keepdim=True).values).sum(dim=1) > 1


_TXT_HEADER = ("iou,pixel_acc,f_score,f_max,f_mean,mae,s_measure,miou_ub,pixel_acc_ub,f_score_ub,f_max_ub,f_mean_ub,"
               "mae_ub,s_measure_ub")


class Evaluator:
    """Same constructor and call signature as the reference Evaluator (evaluator.pyc@L18-32, @L164-174).

    `dataset` may be passed instead of relying on the reference's dataset layer (out of scope, SURVEY.md
    §2.1 #12): any iterable of dicts {'x': b x 3 x H x W float32, 'm': b x 1 x H x W {0,1}} or a callable
    `(dataset_name, batch_size) -> iterable`.
    """

    def __init__(self, network: Callable, arch: str = "vit_small", dir_dataset: Optional[str] = None, visualizer=None,
                 debug: bool = False, dataset=None):
        if dir_dataset is not None:
            assert os.path.exists(dir_dataset), f"{dir_dataset} does not exist"
        self.model, self.arch, self.dir_dataset, self.visualizer, self.debug = network, arch, dir_dataset, visualizer, debug
        self.dataset = dataset
        # When torch.distributed is initialised with more than one rank, a sweep is treated as rank-sharded: every rank evaluates
        # its own batches and the per-image records are gathered before the averages are formed.  Set False to evaluate the same
        # data independently on every rank.
        self.sharded = True
        self.global_records = None   # sharded sweeps: the gathered records of all ranks (device tensors)
        self.objectness_ties = 0     # images of the last sweep whose objectness top-1 was an exact tie (see _count_ties)
        self._ties = []
        self._recs = None            # per-batch device records of the last sweep
        self._records_host = None
        self._copy_stream = None
        self._pinned = None          # page-locked landing buffer for the finalised values (the only per-sweep read-back)

    def _count_ties(self) -> int:
        return int(torch.cat(self._ties).sum().item()) if self._ties else 0

    def tie_flags(self) -> np.ndarray:
        """bool per image of the last sweep (this rank's shard, dataset order): exact objectness top-1 tie."""
        return torch.cat(self._ties).cpu().numpy() if self._ties else np.zeros(0, bool)

    def device_records(self) -> Dict[str, torch.Tensor]:
        """Integer / float64 records of the last sweep, in dataset order, as device tensors (what a multi-GPU caller all-reduces)."""
        if not self._recs:
            raise _lib.SmkError("no sweep has been evaluated yet")
        return {k: torch.cat([getattr(r, k) for r in self._recs]) for k in ("m_counts", "m_sums", "idx", "q_counts")}

    @property
    def records(self) -> Optional[Dict[str, np.ndarray]]:
        """The same records on the host (numpy), fetched on first access: the sweep itself only reads back the finalised values."""
        if self._records_host is None and self._recs:
            self._records_host = {k: v.cpu().numpy() for k, v in self.device_records().items()}
        return self._records_host

    def _batches(self, dataset_name: str, batch_size: int) -> Iterable[dict]:
        if self.dataset is None:
            raise _lib.SmkError("pass dataset=... (an iterable of {'x','m'} batches); the reference's directory readers "
                                "are outside the B200 hot path")
        return self.dataset(dataset_name, batch_size) if callable(self.dataset) else self.dataset

    @torch.no_grad()
    def __call__(self, dataset_name: str, dir_ckpt: Optional[str] = None, img_size: Optional[int] = None, scale_factor: int = 2,
                 batch_size: int = 1, device: torch.device = torch.device("cuda:0"), cost_type: str = "iou") -> dict:
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.SmkError("the B200 evaluator runs on CUDA devices only")
        # Two overlapped streams per batch (the reference does everything serially, with >= 16 syncs per image):
        #   copy stream   : host → device of batch i+1 (pinned host memory makes it asynchronous)
        #   compute stream: model forward, fused evaluation and metric finalisation of batch i (all on the device)
        # The host sees one device → host copy at the end of the sweep: 8 doubles per evaluated mask (+ the integer records).
        compute = torch.cuda.current_stream(device)
        if self._copy_stream is None or self._copy_stream.device != device:
            self._copy_stream = torch.cuda.Stream(device)
        copy = self._copy_stream

        def stage(dict_data):
            x, gt = dict_data["x"], dict_data["m"]
            if gt.device.type == "cpu" and gt.dtype != torch.uint8:
                gt = (gt != 0).to(torch.uint8)                 # 1 byte per pixel over PCIe instead of the loader's int64
            copy.wait_stream(compute)                          # do not run ahead of the batch that still owns the buffers
            with torch.cuda.stream(copy):
                xd, gd = x.to(device, non_blocking=True), gt.to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return xd, gd, ev

        recs, vals = [], []
        self._ties = []
        with torch.cuda.device(device):
            it = iter(self._batches(dataset_name, batch_size))
            first = next(it, None)
            nxt = stage(first) if first is not None else None
            while nxt is not None:
                x, gt, ev = nxt
                following = None if self.debug else next(it, None)
                nxt = stage(following) if following is not None else None
                compute.wait_event(ev)
                x.record_stream(compute)
                gt.record_stream(compute)
                out = self.model(x, encoder_only=False, skip_decoder=False)     # BaseStructure._forward contract
                rec = eval_batch(out["mask_pred"], out["objectness"], gt, up=4)
                self._ties.append(objectness_top1_ties(out["objectness"]))
                recs.append(rec)
                vals.append(finalize_device(rec.m_counts, rec.m_sums))
            import torch.distributed as dist
            sharded = self.sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if not recs and not sharded:
                raise _lib.SmkError("empty dataset")
            self._recs, self._records_host = recs, None
            if sharded:
                # rank-sharded sweep: the one collective of the path gathers every rank's per-image rows (an empty
                # shard contributes none), then every rank finalises the whole sweep and forms the same ordered means
                from .parallel import gather_records
                lc = torch.cat([r.m_counts for r in recs]) if recs else torch.zeros(0, 2, _lib.MCOUNT_STRIDE, dtype=torch.int32, device=device)
                ls = torch.cat([r.m_sums for r in recs]) if recs else torch.zeros(0, 2, _lib.MSUM_STRIDE, dtype=torch.float64, device=device)
                fc, fs = gather_records(lc, ls)
                if fc.shape[0] == 0:
                    raise _lib.SmkError("empty dataset")
                self.global_records = {"m_counts": fc, "m_sums": fs}
                vals = [finalize_device(fc, fs)]
            # one synchronising read-back for the whole sweep: 8 doubles per evaluated mask, into page-locked memory
            dv = torch.cat(vals)
            if self._pinned is None or self._pinned.numel() < dv.numel():
                self._pinned = torch.empty(dv.numel(), dtype=dv.dtype).pin_memory()
            hv = self._pinned[:dv.numel()].view(dv.shape)
            hv.copy_(dv, non_blocking=True)
            compute.synchronize()
            v = values_from_device(hv.numpy().copy())
        res = {k: running_mean(v[k][:, 0]) for k in METRIC_KEYS}
        res.update({k + "_ub": running_mean(v[k][:, 1]) for k in METRIC_KEYS})
        self.objectness_ties = self._count_ties()
        write = dir_ckpt is not None and not (sharded and dist.get_rank() != 0)      # sharded sweeps: rank 0 writes the file
        if write:
            os.makedirs(dir_ckpt, exist_ok=True)
            with open(f"{dir_ckpt}/metrics_{dataset_name}.txt", "w") as f:
                f.write(_TXT_HEADER + "\n")
                f.write(",".join(str(res[k]) for k in METRIC_KEYS) + "," + ",".join(str(res[k + "_ub"]) for k in METRIC_KEYS))
        return res
