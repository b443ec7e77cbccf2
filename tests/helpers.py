"""Test helpers: numpy emulation of the GPU record layout (include/selfmask_b200.h) so that the host-side
finalisation can be checked on CPU against the oracle / golden fixtures."""
import numpy as np

from oracle import selfmask_oracle as O


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
