"""Metric finalisation on the host + drop-in replacements for the reference's five metric callables.

The GPU produces exact integer counts / 256-bin histograms and double-precision moment sums per image
(`smk_eval_batch`, `smk_mask_metrics`); this module turns them into the reference's float32 values with the
same operation order as `metrics/iou.py:22-31`, `f_measure.py:24-81`, `mae.py:9`, `pixel_acc.py:10-14`,
`s_measure.py:11-124` and averages them like `metrics/average_meter.py:12-16`.  Pure numpy, vectorised
over images; no pixel ever reaches the host.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

F32 = np.float32
EPS = F32(1e-7)
METRIC_KEYS = ("iou", "pixel_accuarcy", "f_score", "f_max", "f_mean", "mae", "s_measure")   # [sic] evaluator.pyc@L294-308


def iou_from_counts(inter, union):
    """iou.py:31 — int64 / (int64 + 1e-7) in float32."""
    return (np.asarray(inter).astype(F32) / (np.asarray(union).astype(F32) + EPS)).astype(F32)


def f_from_counts(tp, tp_fp, tp_fn, beta_square: float = 0.3):
    """f_measure.py:24-50; the reference squares beta_square once more (:49,80) → weight 0.09."""
    tp = np.asarray(tp).astype(F32)
    prec = tp / (np.asarray(tp_fp).astype(F32) + EPS)
    rec = tp / (np.asarray(tp_fn).astype(F32) + EPS)
    b2 = beta_square ** 2
    return ((F32(1 + b2) * prec * rec) / (F32(b2) * prec + rec + EPS)).astype(F32)


def counts_above_thresholds(hist: np.ndarray) -> np.ndarray:
    """hist[..., 256] over bin(p) = #{k : t_k < p}  →  count(p > t_k) for k = 0..254 = Σ_{b>k} hist[b].
    Stays in the input's integer width (per-image pixel counts fit int32)."""
    cs = np.cumsum(hist, axis=-1, dtype=hist.dtype if hist.dtype in (np.int32, np.int64) else np.int64)
    return cs[..., 255:256] - cs[..., :255]


def s_measure_from_sums(counts: np.ndarray, sums: np.ndarray, alpha: float = 0.5) -> np.ndarray:
    """s_measure.py:108-124 from moment sums (float64).  counts [...,528] int, sums [...,32] float64.
    Vectorised over records; every expression keeps the operation order of the scalar twin `tests/helpers.py::s_measure_from_sums_loop` (same IEEE
    double results, NaN cases included)."""
    counts = np.asarray(counts)
    sums = np.asarray(sums, np.float64)
    shape = counts.shape[:-1]
    c = counts.reshape(-1, counts.shape[-1])
    s = sums.reshape(-1, sums.shape[-1])
    n, G = c[:, 519].astype(np.float64), c[:, 514].astype(np.float64)
    with np.errstate(all="ignore"):
        mean_p = s[:, 0] / n

        def obj(sum1, sum2, cnt):                          # s_measure.py:54-60, unbiased std
            mu = sum1 / cnt
            var = np.where(cnt > 1, (sum2 - cnt * mu * mu) / (cnt - 1), np.nan)
            sd = np.sqrt(np.maximum(var, 0.0))
            return 2.0 * mu / (mu * mu + 1.0 + sd + 1e-20)
        u = G / n
        s_obj = u * obj(s[:, 3], s[:, 4], G) + (1 - u) * obj(s[:, 5], s[:, 6], n - G)
        X, Y = c[:, 517].astype(F32), c[:, 518].astype(F32)
        hw = n.astype(F32)
        q = s[:, 8:28].reshape(-1, 4, 5)
        w1 = (X * Y) / hw
        w2 = q[:, 1, 0].astype(F32) / hw
        w3 = q[:, 2, 0].astype(F32) / hw
        w4 = F32(1) - w1 - w2 - w3
        Q = []
        for k in range(4):
            N, sp, sp2, sg, spg = (q[:, k, j] for j in range(5))
            x, y = sp / N, sg / N
            den = N - 1 + 1e-20
            sx2 = (sp2 - N * x * x) / den
            sy2 = (sg - N * y * y) / den
            sxy = (spg - N * x * y) / den
            a = 4 * x * y * sxy
            b = (x * x + y * y) * (sx2 + sy2)
            a = np.where(((y == 0.0) | (y == 1.0)) & ~np.isnan(a), 0.0, a)   # (g - ȳ) ≡ 0 → the reference's σxy is an exact zero
            Qk = np.where(a != 0, a / (b + 1e-20), np.where(b == 0, 1.0, 0.0))
            Q.append(np.where(N == 0, np.nan, Qk))
        s_reg = w1.astype(np.float64) * Q[0] + w2.astype(np.float64) * Q[1] + w3.astype(np.float64) * Q[2] + w4.astype(np.float64) * Q[3]
        val = alpha * s_obj + (1 - alpha) * s_reg
        val = np.where(val < 0, 0.0, val)
        out = np.where(G == 0, 1.0 - mean_p, np.where(G == n, mean_p, val))
    return out.reshape(shape)


def finalize(m_counts: np.ndarray, m_sums: np.ndarray) -> Dict[str, np.ndarray]:
    """Per-mask metric values from the GPU records.  m_counts [...,528] int32, m_sums [...,32] float64.
    Returns float32 arrays of shape [...] (s_measure float64, like the reference's python float)."""
    c = np.asarray(m_counts)
    if c.dtype not in (np.int32, np.int64):
        c = c.astype(np.int64)
    s = np.asarray(m_sums, np.float64)
    tp05, tpfp05, G, tpm, tpfpm, n = (c[..., i].astype(np.int64) for i in (512, 513, 514, 515, 516, 519))
    union = tpfp05 + G - tp05
    fg_above = counts_above_thresholds(c[..., 0:256])
    all_above = fg_above + counts_above_thresholds(c[..., 256:512])
    f_k = f_from_counts(fg_above, all_above, G[..., None])
    wrong = tpfp05 + G - 2 * tp05
    return {
        "iou": iou_from_counts(tp05, union),
        "pixel_accuarcy": ((n - wrong).astype(F32) / n.astype(F32)).astype(F32),
        "f_score": f_from_counts(tp05, tpfp05, G),
        "f_max": f_k.max(axis=-1),
        "f_mean": f_from_counts(tpm, tpfpm, G),
        "mae": (s[..., 1] / n).astype(F32),
        "s_measure": s_measure_from_sums(c, s),
    }


def finalize_device(m_counts: torch.Tensor, m_sums: torch.Tensor) -> torch.Tensor:
    """`finalize` on the GPU (smk_finalize_records): device records [..., 528] int32 / [..., 32] float64 → float64
    tensor [..., 8] = the METRIC_KEYS values in order (the six float32 metrics as exact float32 values) + one pad."""
    _lib.require_cuda(m_counts, "m_counts", torch.int32)
    _lib.require_cuda(m_sums, "m_sums", torch.float64)
    m_counts, m_sums = m_counts.contiguous(), m_sums.contiguous()
    n = m_counts.numel() // _lib.MCOUNT_STRIDE
    out = torch.empty(*m_counts.shape[:-1], 8, dtype=torch.float64, device=m_counts.device)
    b2 = 0.3 ** 2
    with torch.cuda.device(m_counts.device):
        check(lib().smk_finalize_records(ptr(m_counts), ptr(m_sums), n, float(F32(1 + b2)), float(F32(b2)), float(EPS), ptr(out),
                                         stream_ptr()), "smk_finalize_records")
    return out


def values_from_device(vals: np.ndarray) -> Dict[str, np.ndarray]:
    """[..., 8] float64 array of `finalize_device` → the dict `finalize` returns (float32 metrics, float64 S-measure)."""
    vals = np.asarray(vals)
    return {k: (vals[..., i] if k == "s_measure" else vals[..., i].astype(F32)) for i, k in enumerate(METRIC_KEYS)}


def running_mean(vals: np.ndarray) -> float:
    """AverageMeter (average_meter.py:12-16): sequential accumulation in the dtype of the values
    (float32 for the tensor-valued metrics, float64 for S-measure's python floats), dataset order."""
    vals = np.asarray(vals)
    if vals.size == 0:
        return 0.0
    return float(np.add.accumulate(vals, dtype=vals.dtype)[-1] / vals.dtype.type(len(vals)))


class AverageMeter:
    """metrics/average_meter.py:1-16, unchanged semantics."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n: int):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


# ---- drop-in metric callables (same names, arguments and return types as metrics/*.py) ---------------

def _device_record(pred_mask: torch.Tensor, gt_mask: torch.Tensor):
    """Run the GPU reduction for [H,W] or [B,H,W] masks; returns (counts [n,528], sums [n,32]) on the host."""
    _lib.require_cuda(pred_mask, "pred_mask")
    if pred_mask.shape != gt_mask.shape:
        raise AssertionError(f"{pred_mask.shape} != {gt_mask.shape}")
    p = pred_mask.detach().to(torch.float32).reshape(-1, *pred_mask.shape[-2:]).contiguous()
    g = (gt_mask.detach().reshape(p.shape) != 0).to(torch.uint8).to(p.device).contiguous()
    n, H, W = p.shape
    counts = torch.empty(n, _lib.MCOUNT_STRIDE, dtype=torch.int32, device=p.device)
    sums = torch.empty(n, _lib.MSUM_STRIDE, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        check(lib().smk_mask_metrics(ptr(p), ptr(g), n, H, W, ptr(counts), ptr(sums), stream_ptr()), "smk_mask_metrics")
    return counts.cpu().numpy(), sums.cpu().numpy()


def _shape_like(pred_mask, arr):
    t = torch.from_numpy(np.asarray(arr))
    return t.reshape(pred_mask.shape[:-2]) if pred_mask.ndim > 2 else t.reshape(())


def compute_iou(pred_mask, gt_mask, threshold: Optional[float] = 0.5, eps: float = 1e-7):
    """metrics/iou.py:6-32 (threshold 0.5, or threshold=None / boolean input for masks that are already binary)."""
    if threshold is None or pred_mask.dtype == torch.bool:
        # the reference uses the mask as it is (any non-zero value is foreground); the GPU reduction thresholds at 0.5, which is
        # the same thing only for {0, 1} masks — anything else is refused instead of silently re-thresholded
        if pred_mask.dtype != torch.bool and not bool(((pred_mask == 0) | (pred_mask == 1)).all()):
            raise _lib.SmkError("compute_iou(threshold=None) expects a binary {0,1} mask on the GPU path")
        pred_mask = pred_mask.to(torch.float32)       # {0,1} > 0.5 reproduces the boolean mask
    elif threshold != 0.5:
        raise _lib.SmkError("only threshold=0.5 is implemented on the GPU path")
    c, s = _device_record(pred_mask, gt_mask)
    return _shape_like(pred_mask, finalize(c, s)["iou"])


class FMeasure:
    """metrics/f_measure.py:4-92."""

    def __init__(self, default_thres: float = 0.5, beta_square: float = 0.3, n_bins: int = 255, eps: float = 1e-7):
        if default_thres != 0.5 or n_bins != 255 or beta_square != 0.3 or eps != 1e-7:
            raise _lib.SmkError("only the reference defaults are implemented on the GPU path")

    def __call__(self, pred_mask: torch.Tensor, gt_mask: torch.Tensor) -> dict:
        """pred_mask, gt_mask: (H x W), as documented by the reference (:83-92); returns 0-d CPU tensors."""
        if pred_mask.ndim != 2:
            raise _lib.SmkError("FMeasure expects (H x W) masks")
        c, s = _device_record(pred_mask, gt_mask)
        f = finalize(c, s)
        return {k: torch.from_numpy(np.asarray(f[src])).reshape(())
                for k, src in (("f_measure", "f_score"), ("f_max", "f_max"), ("f_mean", "f_mean"))}


def compute_mae(pred_mask: torch.Tensor, gt_mask: torch.Tensor) -> torch.Tensor:
    """metrics/mae.py:4-9."""
    c, s = _device_record(pred_mask, gt_mask)
    return _shape_like(pred_mask, finalize(c, s)["mae"])


def compute_pixel_accuracy(pred_mask: torch.Tensor, gt_mask: torch.Tensor, threshold: Optional[float] = 0.5) -> torch.Tensor:
    """metrics/pixel_acc.py:5-14 (threshold 0.5, or threshold=None for masks that are already binary)."""
    if threshold is None:
        if pred_mask.dtype != torch.bool and not bool(((pred_mask == 0) | (pred_mask == 1)).all()):
            raise _lib.SmkError("compute_pixel_accuracy(threshold=None) expects a binary {0,1} mask on the GPU path")
        pred_mask = pred_mask.to(torch.float32)
    elif threshold != 0.5:
        raise _lib.SmkError("only threshold=0.5 is implemented on the GPU path")
    c, s = _device_record(pred_mask, gt_mask)
    return _shape_like(pred_mask, finalize(c, s)["pixel_accuarcy"])


class SMeasure:
    """metrics/s_measure.py:6-124 (returns a python float; unlike the reference it does not mutate gt_mask)."""

    def __init__(self, alpha: float = 0.5):
        self.alpha = alpha

    def __call__(self, pred_mask: torch.Tensor, gt_mask: torch.Tensor) -> float:
        assert pred_mask.shape == gt_mask.shape
        c, s = _device_record(pred_mask, gt_mask >= 0.5)
        return float(s_measure_from_sums(c, s, self.alpha).reshape(-1)[0])
