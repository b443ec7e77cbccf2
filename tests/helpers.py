"""Test helpers: numpy emulation of the GPU record layout (include/selfmask_b200.h) so that the host-side
finalisation can be checked on CPU against the oracle / golden fixtures."""
import math

import numpy as np

from oracle import selfmask_oracle as O

F32 = np.float32


def numpy_record(pred: np.ndarray, gt: np.ndarray):
    """What smk_mask_metrics writes for one full-resolution mask: (counts int32[528], sums float64[32])."""
    p = np.asarray(pred, np.float32)
    g = np.asarray(gt).astype(bool)
    H, W = p.shape
    thr = O.fmax_thresholds()
    bins = np.searchsorted(thr, p.reshape(-1), side="left").reshape(p.shape)     # #{k : t_k < p}
    c = np.zeros(528, np.int64)
    c[0:256] = np.bincount(bins[g], minlength=256)
    c[256:512] = np.bincount(bins[~g], minlength=256)
    b = p > np.float32(0.5)
    c[512], c[513], c[514] = (b & g).sum(), b.sum(), g.sum()
    pd = p.astype(np.float64)
    sp = pd.sum()
    tau = np.float32(2) * np.float32(sp / p.size)
    bm = p > tau
    c[515], c[516] = (bm & g).sum(), bm.sum()
    ng = int(g.sum())
    if ng == 0:
        X, Y = int(round(W / 2)), int(round(H / 2))
    else:
        xs, ys = np.arange(W), np.arange(H)
        X = int(np.rint(np.float32((g.sum(0) * xs).sum()) / np.float32(ng)))
        Y = int(np.rint(np.float32((g.sum(1) * ys).sum()) / np.float32(ng)))
    c[517], c[518], c[519] = X, Y, p.size
    s = np.zeros(32, np.float64)
    s[0], s[1], s[2] = sp, np.abs(pd - g).sum(), float(tau)
    s[3], s[4] = pd[g].sum(), (pd[g] ** 2).sum()
    s[5], s[6] = (1 - pd[~g]).sum(), ((1 - pd[~g]) ** 2).sum()
    quads = [(slice(0, Y), slice(0, X)), (slice(0, Y), slice(X, W)), (slice(Y, H), slice(0, X)), (slice(Y, H), slice(X, W))]
    for k, (ys_, xs_) in enumerate(quads):
        qp, qg = pd[ys_, xs_], g[ys_, xs_]
        s[8 + 5 * k: 13 + 5 * k] = [qp.size, qp.sum(), (qp ** 2).sum(), qg.sum(), qp[qg].sum()]
    return c.astype(np.int32), s


def s_measure_from_sums_loop(counts: np.ndarray, sums: np.ndarray, alpha: float = 0.5) -> np.ndarray:
    """Per-record scalar restatement (python floats) of `selfmask_b200.metrics.s_measure_from_sums` (s_measure.py:108-124 from moment
    sums): the readable definition and the test oracle of the vectorised product version."""
    counts = np.asarray(counts)
    sums = np.asarray(sums, np.float64)
    out = np.empty(counts.shape[:-1], np.float64)
    flat_c, flat_s, flat_o = counts.reshape(-1, counts.shape[-1]), sums.reshape(-1, sums.shape[-1]), out.reshape(-1)
    for i in range(flat_c.shape[0]):
        c, s = flat_c[i], flat_s[i]
        n, G = float(c[519]), float(c[514])
        mean_p = s[0] / n
        if G == 0:
            flat_o[i] = 1.0 - mean_p
            continue
        if G == n:
            flat_o[i] = mean_p
            continue
        with np.errstate(all="ignore"):
            def obj(sum1, sum2, cnt):                      # s_measure.py:54-60, unbiased std
                mu = sum1 / cnt
                var = (sum2 - cnt * mu * mu) / (cnt - 1) if cnt > 1 else float("nan")
                sd = math.sqrt(max(var, 0.0)) if not math.isnan(var) else float("nan")
                return 2.0 * mu / (mu * mu + 1.0 + sd + 1e-20)
            u = G / n
            s_obj = u * obj(s[3], s[4], G) + (1 - u) * obj(s[5], s[6], n - G)
            X, Y = float(c[517]), float(c[518])
            hw = n
            # widths are not in the record: recover W, H from quadrant pixel counts is unnecessary — the
            # weights only need X*Y/area etc., and (W-X)*Y = N_RT, X*(H-Y) = N_LB
            q = s[8:28].reshape(4, 5)
            w1 = F32(F32(X) * F32(Y)) / F32(hw)
            w2 = F32(q[1, 0]) / F32(hw)
            w3 = F32(q[2, 0]) / F32(hw)
            w4 = F32(1) - w1 - w2 - w3
            Q = []
            for k in range(4):
                N, sp, sp2, sg, spg = q[k]
                if N == 0:
                    Q.append(float("nan"))
                    continue
                x, y = sp / N, sg / N
                den = N - 1 + 1e-20
                sx2 = (sp2 - N * x * x) / den
                sy2 = (sg - N * y * y) / den
                sxy = (spg - N * x * y) / den
                a = 4 * x * y * sxy
                b = (x * x + y * y) * (sx2 + sy2)
                if (y == 0.0 or y == 1.0) and not math.isnan(a):
                    a = 0.0                                  # (g - ȳ) ≡ 0 → the reference's σxy is an exact zero
                if a != 0:
                    Q.append(a / (b + 1e-20))
                elif a == 0 and b == 0:
                    Q.append(1.0)
                else:
                    Q.append(0.0)
            s_reg = float(w1) * Q[0] + float(w2) * Q[1] + float(w3) * Q[2] + float(w4) * Q[3]
            val = alpha * s_obj + (1 - alpha) * s_reg
        flat_o[i] = 0.0 if val < 0 else val
    return out
