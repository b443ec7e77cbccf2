"""GPU parity tests of single kernels through the C-ABI: CUDA path vs the CPU oracle / golden fixtures."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import selfmask_b200 as S  # noqa: E402
from oracle import selfmask_oracle as O  # noqa: E402
from selfmask_b200 import metrics as M  # noqa: E402
from selfmask_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from tests.gpu_util import DEV, dev, gemm_bf16, gemm_f32  # noqa: E402
from tests.helpers import numpy_record  # noqa: E402


def test_upsample_bit_exact_with_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "upsample.npz"))
    for src, ref in (("probs", "probs4"), ("odd", "odd4")):
        x = dev(g[src][0])
        n, h, w = x.shape
        out = torch.empty(n, h * 4, w * 4, dtype=torch.float32, device=DEV)
        check(lib().smk_upsample_bilinear(ptr(x), ptr(out), n, h, w, 4, h * 4, w * 4, stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), g[ref][0]), src
    # crop [..., :H, :W] (evaluator.pyc@L211) and scale 1 (identity)
    x = dev(g["odd"][0])
    out = torch.empty(2, 200, 180, dtype=torch.float32, device=DEV)
    check(lib().smk_upsample_bilinear(ptr(x), ptr(out), 2, 52, 48, 4, 200, 180, stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), g["odd4"][0][:, :200, :180])
    out1 = torch.empty_like(x)
    check(lib().smk_upsample_bilinear(ptr(x), ptr(out1), 2, 52, 48, 1, 52, 48, stream_ptr()))
    assert torch.equal(out1, x)


def test_mask_metrics_counts_bit_exact_and_values_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    preds, gts = g["preds"], g["gts"]
    n, H, W = preds.shape
    counts = torch.empty(n, 528, dtype=torch.int32, device=DEV)
    sums = torch.empty(n, 32, dtype=torch.float64, device=DEV)
    d_pred, d_gt = dev(preds), dev(gts.astype(np.uint8))     # keep references: the call is asynchronous
    check(lib().smk_mask_metrics(ptr(d_pred), ptr(d_gt), n, H, W, ptr(counts), ptr(sums), stream_ptr()))
    counts, sums = counts.cpu().numpy(), sums.cpu().numpy()
    for i in range(n):
        c_ref, s_ref = numpy_record(preds[i], gts[i])
        assert np.array_equal(counts[i, :515], c_ref[:515]), i           # histograms, counts@0.5, sum(gt): bit-exact
        assert np.array_equal(counts[i, 517:520], c_ref[517:520]), i     # centroid, pixel count
        assert abs(float(sums[i, 2]) - float(s_ref[2])) <= 1.2e-7, i      # tau = 2*mean(p) (fp32)
        if sums[i, 2] == s_ref[2]:
            assert np.array_equal(counts[i, 515:517], c_ref[515:517]), i
        np.testing.assert_allclose(sums[i], s_ref, rtol=1e-12, atol=1e-9)
    f = M.finalize(counts, sums)
    assert np.array_equal(f["iou"], g["iou"])
    assert np.array_equal(f["f_score"], g["f_measure"])
    assert np.array_equal(f["f_max"], g["f_max"])
    assert np.array_equal(f["pixel_accuarcy"], g["pixel_acc"])
    np.testing.assert_allclose(f["mae"], g["mae"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(f["f_mean"], g["f_mean"], rtol=0, atol=1e-6)
    a, b = f["s_measure"], g["s_measure"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=0, atol=2e-5)


def test_metric_callables_are_drop_in(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for i in (0, 3, 7, 8):
        p, gt = dev(g["preds"][i]), dev(g["gts"][i].astype(np.int64))
        assert S.compute_iou(p, gt).numpy() == g["iou"][i]
        f = S.FMeasure()(p, gt)
        assert f["f_measure"].numpy() == g["f_measure"][i] and f["f_max"].numpy() == g["f_max"][i]
        assert abs(float(f["f_mean"]) - float(g["f_mean"][i])) <= 1e-6
        assert abs(float(S.compute_mae(p, gt)) - float(g["mae"][i])) <= 2e-6
        assert S.compute_pixel_accuracy(p, gt).numpy() == g["pixel_acc"][i]
        assert abs(S.SMeasure()(pred_mask=p, gt_mask=gt.float()) - float(g["s_measure"][i])) <= 2e-5


@pytest.mark.parametrize("B,nq,hp,wp,H,W", [(3, 20, 56, 56, 224, 224), (2, 10, 56, 56, 224, 224), (1, 20, 52, 48, 200, 180),
                                            (1, 20, 96, 96, 384, 384), (1, 3, 5, 7, 20, 28)])
def test_eval_batch_counts_bit_exact_vs_oracle(B, nq, hp, wp, H, W):
    rng = np.random.default_rng(B * 1000 + nq + hp)
    yy, xx = np.mgrid[0:hp, 0:wp].astype(np.float32)
    logits = rng.normal(0, 4, (B, nq, hp, wp)).astype(np.float32)
    for b in range(B):
        for q in range(nq):
            logits[b, q] += 12 * np.exp(-((yy - rng.uniform(0, hp)) ** 2 + (xx - rng.uniform(0, wp)) ** 2) / (2 * rng.uniform(3, 15) ** 2)) - 5
    probs = (1 / (1 + np.exp(-logits))).astype(np.float32)
    obj = rng.random((B, nq)).astype(np.float32)
    gt = O.synth_gt(B, H, W, seed=B + nq, edge_every=2 if B > 1 else 0)
    rec = S.eval_batch(dev(probs), dev(obj), dev(gt), up=4)
    torch.cuda.synchronize()
    qc, idx = rec.q_counts.cpu().numpy(), rec.idx.cpu().numpy()
    mc, ms = rec.m_counts.cpu().numpy(), rec.m_sums.cpu().numpy()
    full = O.upsample_bilinear(probs, 4)[..., :H, :W]
    for b in range(B):
        inter, union = O.iou_counts(full[b], np.broadcast_to(gt[b, 0], full[b].shape))
        assert np.array_equal(qc[b, :, 0], inter) and np.array_equal(qc[b, :, 1], union), b
        sel, ub = int(np.argmax(obj[b])), int(np.argmax(O.iou_from_counts(inter, union)))
        assert tuple(idx[b]) == (sel, ub), (b, idx[b], sel, ub)
        for j, q in enumerate((sel, ub)):
            c_ref, s_ref = numpy_record(full[b, q], gt[b, 0])
            assert np.array_equal(mc[b, j, :515], c_ref[:515]), (b, j)
            assert np.array_equal(mc[b, j, 517:520], c_ref[517:520]), (b, j)
            assert mc[b, j, 520] == q
            if ms[b, j, 2] == s_ref[2]:
                assert np.array_equal(mc[b, j, 515:517], c_ref[515:517]), (b, j)
            np.testing.assert_allclose(ms[b, j], s_ref, rtol=1e-12, atol=1e-9)
    # and the metric values against the oracle's direct (non-histogram) restatement
    vals = M.finalize(mc, ms)
    for b in range(B):
        for j in range(2):
            m = O.image_metrics(full[b, int(mc[b, j, 520])], gt[b, 0])
            for k in ("iou", "f_score", "f_max", "pixel_accuarcy"):
                assert vals[k][b, j] == m[k], (k, b, j)
            assert abs(float(vals["mae"][b, j]) - float(m["mae"])) <= 2e-6
            sa, sb = float(vals["s_measure"][b, j]), m["s_measure"]
            assert (np.isnan(sa) and np.isnan(sb)) or abs(sa - sb) <= 2e-5


@pytest.mark.parametrize("B,nq,hp,wp,H,W", [(2, 20, 56, 56, 224, 224), (2, 7, 52, 48, 200, 180), (1, 5, 96, 96, 384, 384), (2, 4, 3, 2, 12, 8),
                                            (1, 9, 56, 56, 222, 220)])
def test_query_iou_cells_bit_exact_on_threshold_hugging_planes(B, nq, hp, wp, H, W):
    """Cell-classified per-query IoU (query_iou_cells_kernel): planes whose samples sit at 0.5, one ulp either side of it, or are
    noisy — so almost every cell is a boundary cell or a constant field at the threshold — with cropped / tiny geometries."""
    rng = np.random.default_rng(hp * 7 + W)
    half = np.float32(0.5)
    vals = np.array([half, np.nextafter(half, np.float32(1)), np.nextafter(half, np.float32(0)), 0.0, 1.0, 0.49, 0.51], np.float32)
    probs = vals[rng.integers(0, len(vals), (B, nq, hp, wp))]
    noisy = rng.random((B, nq, hp, wp)).astype(np.float32)
    use_noise = rng.random((B, nq, 1, 1)) < 0.3
    probs = np.where(use_noise, noisy, probs).astype(np.float32)
    probs[:, 0] = half                                   # a whole plane exactly at the threshold: nothing is predicted
    probs[:, -1] = np.nextafter(half, np.float32(1))     # ... and one ulp above it: everything is
    obj = rng.random((B, nq)).astype(np.float32)
    gt = (rng.random((B, 1, H, W)) < 0.4).astype(np.uint8)
    rec = S.eval_batch(dev(probs), dev(obj), dev(gt), up=4)
    torch.cuda.synchronize()
    qc = rec.q_counts.cpu().numpy()
    full = O.upsample_bilinear(probs, 4)[..., :H, :W]
    for b in range(B):
        inter, union = O.iou_counts(full[b], np.broadcast_to(gt[b, 0], full[b].shape))
        assert np.array_equal(qc[b, :, 0], inter) and np.array_equal(qc[b, :, 1], union), b
        assert qc[b, 0, 0] == 0 and qc[b, -1, 1] == H * W


def test_device_finalisation_is_bit_identical_to_host(golden_dir):
    """smk_finalize_records == metrics.finalize (numpy) on real records (golden masks, edge-case GTs incl. the reference's
    NaN S-measure cases) and on randomised records: every float32 metric and the float64 S-measure, bit for bit."""
    from selfmask_b200 import metrics as M
    from tests.helpers import numpy_record
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    recs = [numpy_record(p, gt) for p, gt in zip(g["preds"], g["gts"])]
    rng = np.random.default_rng(5)
    H, W = 40, 56
    for kind in range(14):
        p = rng.random((H, W), dtype=np.float32)
        gt = np.zeros((H, W), bool)
        if kind == 1: gt[:] = True
        elif kind == 2: gt[0, 0] = True
        elif kind == 3: gt[:, 0] = True
        elif kind == 4: gt[0, :] = True
        elif kind == 5: gt[H - 1, W - 1] = True
        elif kind == 6: gt[10:20, 10:30] = True; p[:] = 0.0
        elif kind == 7: gt[10:20, 10:30] = True; p[:] = 1.0
        elif kind == 8: gt[10:20, 10:30] = True; p = gt.astype(np.float32)
        elif kind >= 9: gt = rng.random((H, W)) > 0.15 * (kind - 8)
        recs.append(numpy_record(p, gt))
    counts, sums = np.stack([r[0] for r in recs]), np.stack([r[1] for r in recs])
    # randomised but self-consistent histogram records
    n_rand = 300
    rc = np.zeros((n_rand, 528), np.int32)
    rs = np.zeros((n_rand, 32), np.float64)
    for i in range(n_rand):
        p = rng.random((24, 32), dtype=np.float32) ** rng.uniform(0.3, 3.0)
        gt = rng.random((24, 32)) < rng.uniform(0.02, 0.98)
        rc[i], rs[i] = numpy_record(p, gt)
    counts, sums = np.concatenate([counts, rc]), np.concatenate([sums, rs])
    host = M.finalize(counts, sums)
    dev = M.values_from_device(M.finalize_device(torch.from_numpy(counts).to(DEV), torch.from_numpy(sums).to(DEV)).cpu().numpy())
    assert np.isnan(host["s_measure"]).any()
    for k in M.METRIC_KEYS:
        assert host[k].dtype == dev[k].dtype, k
        assert np.array_equal(host[k], dev[k], equal_nan=True), (k, np.nonzero(~((host[k] == dev[k]) | (np.isnan(host[k]) & np.isnan(dev[k]))))[0][:5])


def test_layernorm_matches_torch():
    torch.manual_seed(0)
    x = torch.randn(1000, 384, device=DEV) * 3 + 1
    g, b = torch.randn(384, device=DEV), torch.randn(384, device=DEV)
    ref = torch.nn.functional.layer_norm(x, (384,), g, b, 1e-6)
    y = torch.empty_like(x)
    check(lib().smk_layernorm(ptr(x), ptr(g), ptr(b), ptr(y), 1000, 384, 1e-6, 0, stream_ptr()))
    assert (y - ref).abs().max().item() <= 2e-5
    yb = torch.empty(1000, 384, dtype=torch.bfloat16, device=DEV)
    check(lib().smk_layernorm(ptr(x), ptr(g), ptr(b), ptr(yb), 1000, 384, 1e-6, 1, stream_ptr()))
    assert torch.equal(yb, y.to(torch.bfloat16)) or (yb.float() - ref).abs().max().item() <= 0.05


@pytest.mark.parametrize("M_,N,K,epi", [(197, 384, 384, 0), (1000, 1152, 384, 0), (333, 1536, 384, 1), (130, 384, 1536, 4),
                                        (64, 128, 768, 2), (5, 384, 384, 0)])
def test_gemm_f32_matches_torch(M_, N, K, epi):
    torch.manual_seed(1)
    A, W, bias = torch.randn(M_, K, device=DEV), torch.randn(N, K, device=DEV) * 0.05, torch.randn(N, device=DEV)
    C0 = torch.randn(M_, N, device=DEV)
    ref = (A.double() @ W.double().t() + bias.double())
    if epi & 1:
        ref = torch.nn.functional.gelu(ref)
    if epi & 2:
        ref = torch.relu(ref)
    if epi & 4:
        ref = ref + C0.double()
    out = gemm_f32(A, W, bias, epi, C_init=C0 if epi & 4 else None)
    assert (out.double() - ref).abs().max().item() <= 2e-4


@pytest.mark.parametrize("M_,N,K,epi,f32", [(128, 128, 64, 0, True), (128, 128, 384, 0, True), (197, 384, 384, 0, False),
                                            (1000, 1152, 384, 0, False), (333, 1536, 384, 1, False), (130, 384, 1536, 4, True),
                                            (50432, 1536, 384, 1, False), (50432, 384, 1536, 4, True), (40000, 4608, 384, 0, False),
                                            (50432, 1152, 384, 0, False), (50432, 384, 384, 4, True), (5120, 768, 1152, 0, False),
                                            (64, 256, 768, 2, True), (300, 256, 192, 0, True), (1000, 384, 128, 0, False),
                                            (20000, 768, 256, 1, False), (19000, 1152, 384, 4, True),
                                            # swap-AB form (narrow fp32 output, K >= 1024, >= 16384 rows): ragged token counts, plain / ReLU / residual
                                            (16500, 384, 1024, 4, True), (17001, 256, 2048, 0, True), (16421, 128, 1024, 2, True),
                                            (25216, 384, 1536, 4, True),
                                            # shapes of the opt-in bf16 swap-AB form (SMK_GEMM_SWAP_AB=1); normal form by default
                                            (16500, 1152, 384, 0, False), (20001, 384, 256, 2, False)])
def test_gemm_bf16_tcgen05_matches_torch(M_, N, K, epi, f32):
    torch.manual_seed(2)
    A = torch.randn(M_, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    C0 = torch.randn(M_, N, device=DEV) if epi & 4 else None
    out = gemm_bf16(A, W, bias, epi, out_f32=f32, C_init=C0)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias          # fp32 accumulate of the same bf16 operands
    if epi & 1:
        ref = torch.nn.functional.gelu(ref)
    if epi & 2:
        ref = torch.relu(ref)
    if epi & 4:
        ref = ref + C0
    err = (out.float() - ref).abs().max().item()
    tol = 2e-3 if f32 else 0.02 * max(1.0, ref.abs().max().item())     # bf16 output rounding: 2^-8 relative
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("is_bf16", [False, True])
@pytest.mark.parametrize("Lq,Lk", [(197, 197), (20, 196), (20, 20), (577, 577)])
def test_attention_matches_torch(Lq, Lk, is_bf16):
    torch.manual_seed(3)
    B, H, dh = 2, 6, 64
    dt = torch.bfloat16 if is_bf16 else torch.float32
    q = torch.randn(B, Lq, H * dh, device=DEV).to(dt)
    k = torch.randn(B, Lk, H * dh, device=DEV).to(dt)
    v = torch.randn(B, Lk, H * dh, device=DEV).to(dt)
    o = torch.empty(B, Lq, H * dh, device=DEV, dtype=dt)
    check(lib().smk_attention(ptr(q), ptr(k), ptr(v), ptr(o), B, H, dh, Lq, Lk, Lq * H * dh, H * dh, Lk * H * dh, H * dh,
                              Lk * H * dh, H * dh, Lq * H * dh, H * dh, 0.125, 1 if is_bf16 else 0, stream_ptr()))
    qh = q.float().view(B, Lq, H, dh).transpose(1, 2)
    kh = k.float().view(B, Lk, H, dh).transpose(1, 2)
    vh = v.float().view(B, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B, Lq, H * dh)
    assert (o.float() - ref).abs().max().item() <= (0.02 if is_bf16 else 2e-5)


@pytest.mark.parametrize("B,N", [(1, 197), (3, 197), (2, 64), (2, 256), (5, 130), (64, 197), (3, 208), (2, 193), (2, 17), (150, 197)])
def test_attention_tcgen05_matches_torch(B, N):
    torch.manual_seed(4)
    H, dh = 6, 64
    D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 1.5).to(torch.bfloat16)
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    check(lib().smk_attention_tc(ptr(qkv), ptr(out), B, N, H, 0.125, stream_ptr()), "smk_attention_tc")
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, N, H, dh).transpose(1, 2) for t in qkv.view(B, N, 3 * D).split(D, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, D)
    err = (out.float() - ref).abs().max().item()
    # bf16 P and a bf16 output of magnitude up to ~4: one output ulp is 0.016-0.03; the maximum over 11 M outputs (B = 150) reaches 0.0303
    assert err <= 0.04, err
    assert (out.float() - ref).abs().mean().item() <= 2e-3


def test_bf16x3_split_gemm_is_near_fp32():
    """Decoder tail: A·W^T through ONE tcgen05 GEMM over K' = 3K with the 3-term bf16 split folded into K."""
    torch.manual_seed(5)
    for M_, N, K, epi in [(5120, 384, 384, 0), (5120, 1536, 384, 2), (400, 384, 1536, 0), (20, 768, 384, 0)]:
        A = torch.randn(M_, K, device=DEV) * 2
        W = torch.randn(N, K, device=DEV) * 0.05
        bias = torch.randn(N, device=DEV)
        a3 = torch.empty(M_, 3 * K, dtype=torch.bfloat16, device=DEV)
        w3 = torch.empty(N, 3 * K, dtype=torch.bfloat16, device=DEV)
        check(lib().smk_split3(ptr(A), M_, K, ptr(a3), 0, stream_ptr()))
        check(lib().smk_split3(ptr(W), N, K, ptr(w3), 1, stream_ptr()))
        out = gemm_bf16(a3, w3, bias, epi, out_f32=True)
        ref = A.double() @ W.double().t() + bias.double()
        if epi & 2:
            ref = torch.relu(ref)
        err = (out.double() - ref).abs().max().item()
        assert err <= 3e-4, (M_, N, K, err)     # plain bf16 operands would give ~3e-2 here


@pytest.mark.parametrize("Lq,Lk,kv_rows,kv_row0", [(20, 196, 197, 1), (20, 20, 20, 0), (10, 196, 197, 1), (128, 250, 256, 3)])
def test_attention_tcgen05_general_layout(Lq, Lk, kv_rows, kv_row0):
    """Decoder cross / self attention layout: separate q, k, v matrices, per-image key offset (cls skipped), fp32 out."""
    torch.manual_seed(6)
    B, H, dh = 5, 6, 64
    D = H * dh
    q = torch.randn(B * Lq, D, device=DEV).to(torch.bfloat16)
    kv = torch.randn(B * kv_rows, 2 * D, device=DEV).to(torch.bfloat16)       # k | v interleaved per row (ld = 2D)
    out = torch.zeros(B * Lq, D, device=DEV, dtype=torch.float32)
    k_view, v_view = kv[:, :D], kv[:, D:]
    check(lib().smk_attention_tc_general(ptr(q), D, C.c_void_p(k_view.data_ptr()), 2 * D, C.c_void_p(v_view.data_ptr()), 2 * D,
                                         B * kv_rows, kv_rows, kv_row0, ptr(out), D, 1, B, Lq, Lk, H, 0.125, stream_ptr()))
    torch.cuda.synchronize()
    qh = q.float().view(B, Lq, H, dh).transpose(1, 2)
    kk = kv.float().view(B, kv_rows, 2 * D)[:, kv_row0:kv_row0 + Lk]
    kh = kk[..., :D].reshape(B, Lk, H, dh).transpose(1, 2)
    vh = kk[..., D:].reshape(B, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B * Lq, D)
    assert (out - ref).abs().max().item() <= 0.03


@pytest.mark.parametrize("Lq,Lk,kv_rows,kv_row0,out_mode", [(20, 196, 197, 1, 1), (20, 20, 20, 0, 1), (10, 196, 197, 1, 0), (10, 10, 10, 0, 2),
                                                             (20, 196, 197, 1, 2), (32, 256, 256, 0, 1), (1, 1, 3, 2, 1), (17, 100, 120, 5, 1),
                                                             (20, 144, 145, 1, 2)])
def test_attention_small_matches_torch(Lq, Lk, kv_rows, kv_row0, out_mode):
    """Few-query decoder attention (mma.sync kernel): separate q / k / v matrices, per-image key offset (cls skipped),
    bf16 / fp32 / bf16x3-split outputs; B*heads is large enough to cover several CTAs per SM."""
    torch.manual_seed(16)
    B, H, dh = 37, 6, 64
    D = H * dh
    q = torch.randn(B * Lq, D, device=DEV).to(torch.bfloat16)
    kv = torch.randn(B * kv_rows, 2 * D, device=DEV).to(torch.bfloat16)       # k | v interleaved per row (ld = 2D)
    k_view, v_view = kv[:, :D], kv[:, D:]
    if out_mode == 1:
        out = torch.full((B * Lq, D), 7.0, device=DEV, dtype=torch.float32)
    else:
        out = torch.full((B * Lq, D * (3 if out_mode == 2 else 1)), 7.0, device=DEV, dtype=torch.bfloat16)
    check(lib().smk_attention_small(ptr(q), D, C.c_void_p(k_view.data_ptr()), 2 * D, C.c_void_p(v_view.data_ptr()), 2 * D,
                                    kv_rows, kv_row0, ptr(out), out.shape[1], out_mode, B, Lq, Lk, H, 0.125, stream_ptr()))
    torch.cuda.synchronize()
    qh = q.float().view(B, Lq, H, dh).transpose(1, 2)
    kk = kv.float().view(B, kv_rows, 2 * D)[:, kv_row0:kv_row0 + Lk]
    kh = kk[..., :D].reshape(B, Lk, H, dh).transpose(1, 2)
    vh = kk[..., D:].reshape(B, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B * Lq, D)
    if out_mode == 2:
        hi, hi2, lo = out[:, :D].float(), out[:, D:2 * D].float(), out[:, 2 * D:].float()
        assert torch.equal(hi, hi2)
        got = hi + lo
    else:
        got = out.float()
    tol = 0.03 if out_mode != 0 else 0.05
    assert (got - ref).abs().max().item() <= tol


def test_gemm_bf16_tcgen05_split3_output():
    """out_f32 = 2: the epilogue writes the bf16x3 split [hi | hi | lo] of the fp32 result (decoder FFN hidden, objectness hidden)."""
    torch.manual_seed(7)
    M_, N, K = 700, 384, 384
    A = torch.randn(M_, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    out = torch.zeros(M_, 3 * N, dtype=torch.bfloat16, device=DEV)
    check(lib().smk_gemm_bf16(ptr(A), K, ptr(W), ptr(bias), ptr(out), 3 * N, M_, N, K, 2, 2, stream_ptr()), "gemm split3")
    torch.cuda.synchronize()
    ref = torch.relu(A.float() @ W.float().t() + bias)
    hi, hi2, lo = out[:, :N].float(), out[:, N:2 * N].float(), out[:, 2 * N:].float()
    assert torch.equal(hi, hi2)
    assert (hi - ref).abs().max().item() <= 0.02 * max(1.0, ref.abs().max().item())
    assert (hi + lo - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())      # hi + lo carries ~16 mantissa bits


def test_attention_tcgen05_split3_output():
    torch.manual_seed(8)
    B, H, dh, Lq, Lk = 5, 6, 64, 20, 196
    D = H * dh
    q = torch.randn(B * Lq, D, device=DEV).to(torch.bfloat16)
    k = torch.randn(B * Lk, D, device=DEV).to(torch.bfloat16)
    v = torch.randn(B * Lk, D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(B * Lq, 3 * D, device=DEV, dtype=torch.bfloat16)
    check(lib().smk_attention_tc_general(ptr(q), D, ptr(k), D, ptr(v), D, B * Lk, Lk, 0, ptr(out), 3 * D, 2, B, Lq, Lk, H, 0.125, stream_ptr()))
    torch.cuda.synchronize()
    qh = q.float().view(B, Lq, H, dh).transpose(1, 2)
    kh = k.float().view(B, Lk, H, dh).transpose(1, 2)
    vh = v.float().view(B, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B * Lq, D)
    hi, hi2, lo = out[:, :D].float(), out[:, D:2 * D].float(), out[:, 2 * D:].float()
    assert torch.equal(hi, hi2)
    assert (hi + lo - ref).abs().max().item() <= 0.03
    assert (lo.abs() <= hi.abs() * 2.0 ** -7 + 1e-30).all()          # lo is the rounding residue of hi


@pytest.mark.parametrize("Lq,Lk,split,out_mode", [(577, 577, False, 0), (785, 785, False, 1), (197, 197, True, 2), (577, 577, True, 2),
                                                   (64, 64, False, 1), (1, 1, True, 1), (130, 67, True, 1)])
def test_attention_fa_matches_torch(Lq, Lk, split, out_mode):
    """Online-softmax mma.sync attention (smk_attn_fa.cu): long sequences (384x384 → 577 tokens, ViT-S/8 → 785) with bf16
    operands, and the bf16x3 split mode (hi + lo operands, 3-term products) which must be ~fp32-accurate."""
    torch.manual_seed(21)
    B, H, dh = 3, 6, 64
    D = H * dh
    q32 = torch.randn(B * Lq, D, device=DEV)
    k32 = torch.randn(B * Lk, D, device=DEV)
    v32 = torch.randn(B * Lk, D, device=DEV)

    def parts(x):
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return hi.contiguous(), lo.contiguous()

    (qh, ql), (kh, kl), (vh, vl) = parts(q32), parts(k32), parts(v32)
    width = D * (3 if out_mode == 2 else 1)
    out = torch.full((B * Lq, width), 7.0, device=DEV, dtype=torch.float32 if out_mode == 1 else torch.bfloat16)
    lo_ptr = (lambda t: ptr(t)) if split else (lambda t: None)
    check(lib().smk_attention_fa(ptr(qh), lo_ptr(ql), D, ptr(kh), lo_ptr(kl), D, ptr(vh), lo_ptr(vl), D, Lq, Lk, 0, ptr(out), width,
                                 out_mode, B, Lq, Lk, H, 0.125, stream_ptr()))
    torch.cuda.synchronize()
    src = (qh.float() + ql.float(), kh.float() + kl.float(), vh.float() + vl.float()) if split else (qh.float(), kh.float(), vh.float())
    qq, kk, vv = (t.double().view(B, -1, H, dh).transpose(1, 2) for t in src)
    ref = (torch.softmax(qq @ kk.transpose(-1, -2) * 0.125, -1) @ vv).transpose(1, 2).reshape(B * Lq, D).float()
    if out_mode == 2:
        hi, hi2, lo = out[:, :D].float(), out[:, D:2 * D].float(), out[:, 2 * D:].float()
        assert torch.equal(hi, hi2)
        got = hi + lo
    else:
        got = out.float()
    err = (got - ref).abs().max().item()
    if split:
        assert err <= (3e-4 if out_mode == 2 else 5e-5), err      # 3-term products: fp32-like; the split output keeps ~16 bits
    else:
        assert err <= (0.03 if out_mode == 1 else 0.05), err


@pytest.mark.parametrize("M_,K", [(300, 384), (50432, 384), (19000, 1536), (128, 64)])
def test_gemm_layernorm_fused_matches_torch(M_, K):
    """smk_gemm_ln: X += A·W^T + bias (fp32 residual stream, in place) and Xn = LayerNorm(X) in bf16 from the same tile."""
    torch.manual_seed(31)
    N = 384
    A = torch.randn(M_, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=DEV), 1 + 0.1 * torch.randn(N, device=DEV), 0.1 * torch.randn(N, device=DEV)
    X0 = torch.randn(M_, N, device=DEV) * 2 + 0.5
    X = X0.clone()
    Xn = torch.full((M_, N), 7.0, device=DEV, dtype=torch.bfloat16)
    check(lib().smk_gemm_ln(ptr(A), K, ptr(W), ptr(bias), ptr(X), ptr(gamma), ptr(beta), ptr(Xn), M_, N, K, 1e-6, stream_ptr()), "gemm_ln")
    torch.cuda.synchronize()
    ref_x = X0 + A.float() @ W.float().t() + bias
    assert (X - ref_x).abs().max().item() <= 2e-3
    ref_n = torch.nn.functional.layer_norm(X, (N,), gamma, beta, 1e-6)       # LayerNorm of the kernel's own fp32 output
    assert (Xn.float() - ref_n).abs().max().item() <= 0.02 * max(1.0, ref_n.abs().max().item())


# ---- fp16s mode kernels --------------------------------------------------------------------------------------------------------
def _split16(x, dt):
    hi = x.to(dt)
    return hi, (x - hi.float()).to(dt)


@pytest.mark.parametrize("f16", [1, 0])
@pytest.mark.parametrize("M_,N,K,epi,terms,out_kind", [
    (50432, 1152, 384, 0, 2, 0),      # qkv: A_hi·(W_hi + W_lo), 16-bit output (swap-AB form: transposed 16-bit epilogue)
    (50432, 384, 384, 4, 3, 1),       # proj: 3 terms, fp32 residual
    (50432, 1536, 384, 1, 3, 3),      # fc1: 3 terms, GELU, [hi | lo] output
    (50432, 384, 1536, 4, 3, 1),      # fc2: 3 terms over K = 1536 (swap-AB form), fp32 residual
    (20000, 4608, 384, 0, 2, 0),      # memory K/V
    (700, 384, 768, 0, 3, 1), (130, 256, 128, 2, 1, 3), (333, 128, 64, 0, 3, 2), (5120, 768, 384, 0, 2, 1)])
def test_gemm_split_terms_match_fp64(M_, N, K, epi, terms, out_kind, f16):
    """smk_gemm_split: operands stored as [hi | lo] rows, 1 / 2 / 3 tensor-core terms selected by column offsets; fp16 and bf16."""
    torch.manual_seed(40)
    dt = torch.float16 if f16 else torch.bfloat16
    A32 = torch.randn(M_, K, device=DEV) * 1.5
    W32 = torch.randn(N, K, device=DEV) * 0.05
    bias = torch.randn(N, device=DEV)
    (ah, al), (wh, wl) = _split16(A32, dt), _split16(W32, dt)
    A2 = torch.cat([ah, al], dim=1).contiguous()
    W2 = torch.cat([wh, wl], dim=1).contiguous()
    a_off, w_off = {1: ([0], [0]), 2: ([0, 0], [0, K]), 3: ([0, 0, K], [0, K, 0])}[terms]
    C0 = torch.randn(M_, N, device=DEV) if epi & 4 else None
    if out_kind == 1:
        out = C0.clone() if C0 is not None else torch.empty(M_, N, device=DEV)
    else:
        out = torch.full((M_, N * {0: 1, 2: 3, 3: 2}[out_kind]), 7.0, dtype=dt, device=DEV)
    ao, wo = (C.c_int32 * 3)(*a_off), (C.c_int32 * 3)(*w_off)
    check(lib().smk_gemm_split(ptr(A2), 2 * K, ptr(W2), 2 * K, ptr(bias), ptr(out), out.shape[1], M_, N, K, epi, out_kind, f16, terms, ao, wo,
                               stream_ptr()), "smk_gemm_split")
    torch.cuda.synchronize()
    # the exact value of the issued terms (fp64), and the full-precision product they approximate
    Ah, Al, Wh, Wl = ah.double(), al.double(), wh.double(), wl.double()
    issued = Ah @ Wh.t()
    if terms >= 2:
        issued = issued + Ah @ Wl.t()
    if terms >= 3:
        issued = issued + Al @ Wh.t()
    issued = issued + bias.double()
    act = (lambda t: torch.nn.functional.gelu(t)) if epi & 1 else ((lambda t: torch.relu(t)) if epi & 2 else (lambda t: t))
    ref = act(issued) + (C0.double() if C0 is not None else 0)
    if out_kind == 1:
        got = out.double()
    elif out_kind == 0:
        got = out.double()
    elif out_kind == 3:
        got = out[:, :N].double() + out[:, N:].double()
    else:
        assert torch.equal(out[:, :N], out[:, N:2 * N])
        got = out[:, :N].double() + out[:, 2 * N:].double()
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    if out_kind == 0:
        assert err <= (2.0 ** -10 if f16 else 2.0 ** -7) * scale, err          # one rounding of the 16-bit output
    else:
        assert err <= 2e-4 * scale, err                                        # fp32 accumulate (+ gelu_fast's 2e-5)
    if terms == 3:      # and the 3-term product is ~fp32-accurate against the unsplit fp32 operands
        full = act(A32.double() @ W32.double().t() + bias.double()) + (C0.double() if C0 is not None else 0)
        assert (got - full).abs().max().item() <= (2e-4 if f16 else 6e-4) * scale


@pytest.mark.parametrize("f16,Lk", [(1, 576), (0, 576), (1, 784), (0, 200)])
def test_attention_tcgen05_multi_tile_decoder_cross_layout(f16, Lk):
    """Decoder cross-attention at 384 x 384 / ViT-S/8 on the multi-key-tile kernel: 20 queries per image against the patch tokens of the
    all-layer K/V tensor (row stride L·2·D, cls row skipped); fp16 operands write bf16 [hi | hi | lo] parts (out_mode | 8)."""
    torch.manual_seed(47)
    B, nq, H, dh, L = 5, 20, 6, 64, 6
    D = H * dh
    N = Lk + 1
    dt = torch.float16 if f16 else torch.bfloat16
    q = (torch.randn(B * nq, D, device=DEV) * 1.5).to(dt)
    kv = (torch.randn(B * N, L * 2 * D, device=DEV) * 1.5).to(dt)
    layer = 3
    kl = kv[:, layer * 2 * D:]
    out = torch.full((B * nq, 3 * D), 7.0, device=DEV, dtype=torch.bfloat16 if f16 else dt)
    check(lib().smk_attention_tc_multi(ptr(q), D, ptr(kl), L * 2 * D, ptr(kl[:, D:]), L * 2 * D, B * nq, B * N, nq, N, 1, ptr(out), 3 * D,
                                       2 | (8 if f16 else 0), B, nq, Lk, H, 0.125, f16, stream_ptr()), "smk_attention_tc_multi")
    torch.cuda.synchronize()
    k = kv.view(B, N, L * 2 * D)[:, 1:, layer * 2 * D:layer * 2 * D + D].double().reshape(B, Lk, H, dh).transpose(1, 2)
    v = kv.view(B, N, L * 2 * D)[:, 1:, layer * 2 * D + D:layer * 2 * D + 2 * D].double().reshape(B, Lk, H, dh).transpose(1, 2)
    qq = q.double().view(B, nq, H, dh).transpose(1, 2)
    ref = (torch.softmax(qq @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * nq, D)
    assert torch.equal(out[:, :D], out[:, D:2 * D])
    got = out[:, :D].double() + out[:, 2 * D:].double()
    assert (got - ref).abs().max().item() <= (5e-3 if f16 else 0.04)


@pytest.mark.parametrize("n_img,hw,f16", [(1, 64, 0), (2, 64, 1), (3, 196, 0), (8, 196, 2), (5, 156, 1), (40, 9, 0), (2, 576, 2), (256, 196, 2)])
def test_gemm_token_assembly(n_img, hw, f16):
    """Patch-embed form: rows re-indexed around the class-token rows (3-D output map, chunks that cross image boundaries are stored once
    per image), position embedding + bias added, class-token rows untouched.  bf16 / fp16 / fp8-corrected operands; hw smaller than a
    32-row chunk (several images per chunk), multiples of 32 (no crossing) and the 224 / 384 geometries."""
    torch.manual_seed(48)
    N, K = 384, 192
    A32 = torch.randn(n_img * hw, K, device=DEV)
    W32 = torch.randn(N, K, device=DEV) * 0.05
    bias = torch.randn(N, device=DEV)
    pos = torch.randn(hw + 1, N, device=DEV)
    if f16 == 2:
        A = torch.zeros(n_img * hw, 2 * K, dtype=torch.float16, device=DEV)
        W = torch.zeros(N, 2 * K, dtype=torch.float16, device=DEV)
        check(lib().smk_split_q8(ptr(A32), K, ptr(A), n_img * hw, K, 0, stream_ptr()))
        check(lib().smk_split_q8(ptr(W32), K, ptr(W), N, K, 1, stream_ptr()))
        ld, Aeff, Weff = 2 * K, A32.double(), W32.double()
    else:
        dt = torch.float16 if f16 else torch.bfloat16
        A, W = A32.to(dt), W32.to(dt)
        ld, Aeff, Weff = K, A.double(), W.double()
    Cc = torch.full((n_img, hw + 1, N), 7.0, device=DEV)
    check(lib().smk_gemm_tokens(ptr(A), ld, ptr(W), ld, ptr(bias), ptr(pos), ptr(Cc), N, n_img, hw, N, K, f16, stream_ptr()), "smk_gemm_tokens")
    torch.cuda.synchronize()
    ref = (Aeff @ Weff.t() + bias.double()).view(n_img, hw, N) + pos[1:].double()[None]
    assert torch.equal(Cc[:, 0], torch.full((n_img, N), 7.0, device=DEV))            # class-token rows untouched
    assert (Cc[:, 1:].double() - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def _q8(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).double()


def _unpack_q8(rows2k, K):
    """[rows, 2K fp16 columns] → (hi fp64 [rows, K], first fp64 [rows, K], second fp64 [rows, K]) of the split_q8 row layout."""
    hi = rows2k[:, :K].double()
    q = rows2k[:, K:].contiguous().view(torch.uint8).view(rows2k.shape[0], K // 32, 2, 32)
    first = q[:, :, 0, :].reshape(-1, K).view(torch.float8_e4m3fn).double()
    second = q[:, :, 1, :].reshape(-1, K).view(torch.float8_e4m3fn).double()
    return hi, first, second


def test_split_q8_layout_and_values():
    """smk_split_q8: fp16 hi + the e4m3 correction operands (activation and weight scalings), 32-column interleave."""
    torch.manual_seed(44)
    rows, K = 37, 192
    x = torch.randn(rows, K, device=DEV) * 3.0
    x[0, :8] = torch.tensor([500.0, -700.0, 1e-6, 0.0, 447.0, -3e-4, 60000.0, 1.0], device=DEV)      # saturation / underflow corners
    for is_w in (0, 1):
        out = torch.zeros(rows, 2 * K, dtype=torch.float16, device=DEV)
        check(lib().smk_split_q8(ptr(x), K, ptr(out), rows, K, is_w, stream_ptr()), "smk_split_q8")
        torch.cuda.synchronize()
        hi, first, second = _unpack_q8(out, K)
        h = x.to(torch.float16)
        lo = (x - h.float())
        assert torch.equal(hi, h.double())
        if is_w:
            assert torch.equal(first, _q8(lo * 2.0 ** 15)) and torch.equal(second, _q8(h.float() * 16.0))
        else:
            assert torch.equal(first, _q8(h.float())) and torch.equal(second, _q8(lo * 2.0 ** 11))


@pytest.mark.parametrize("rows", [1, 37, 50432])
def test_layernorm_fp16_operand_rows(rows):
    """The fp16s LayerNorm writes its consumer GEMM's operand rows in the same pass as the fp32 result: [hi | lo] fp16 and [hi | e4m3]
    must be exactly the splits of that fp32 row (bit for bit), and the fp32 row the reference LayerNorm within rounding."""
    torch.manual_seed(49)
    D = 384
    x = torch.randn(rows, D, device=DEV) * 3.0 + 0.5
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y32 = torch.empty(rows, D, device=DEV)
    for lo_kind in (0, 1, 2):
        y = torch.full((rows, 2 * D), 7.0, dtype=torch.float16, device=DEV)
        check(lib().smk_layernorm_f16(ptr(x), ptr(g), ptr(b), ptr(y), 2 * D, ptr(y32), rows, D, 1e-6, lo_kind, stream_ptr()), "smk_layernorm_f16")
        torch.cuda.synchronize()
        hi = y32.to(torch.float16)
        assert torch.equal(y[:, :D], hi)
        if lo_kind == 0:
            assert (y[:, D:] == 7.0).all()
        elif lo_kind == 1:
            assert torch.equal(y[:, D:], (y32 - hi.float()).to(torch.float16))
        else:
            want = torch.zeros(rows, 2 * D, dtype=torch.float16, device=DEV)
            check(lib().smk_split_q8(ptr(y32), D, ptr(want), rows, D, 0, stream_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(y.view(torch.int16), want.view(torch.int16))
    ref = torch.nn.functional.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-6)
    assert (y32.double() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("B,H,W,P,is_u8", [(3, 224, 224, 16, 1), (2, 200, 180, 16, 0), (2, 224, 224, 8, 1), (1, 40, 56, 16, 1)])
def test_im2col_fp16_operand_rows(B, H, W, P, is_u8):
    """Patch im2col of the fp16s mode: the [hi | e4m3] rows are the smk_split_q8 form of the [hi | lo] rows' values (hi + lo carries the
    normalised pixel to ~2^-22), uint8 pixels normalised with the reference's IEEE expression, zero padding to multiples of P."""
    torch.manual_seed(50)
    mean_std = (C.c_float * 6)(0.485, 0.456, 0.406, 0.229, 0.224, 0.225)
    if is_u8:
        x = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device=DEV)
        xn = ((x.float() / 255.0) - torch.tensor([0.485, 0.456, 0.406], device=DEV).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225], device=DEV).view(1, 3, 1, 1)
    else:
        x = torch.randn(B, 3, H, W, device=DEV)
        xn = x
    hp, wp, K = -(-H // P), -(-W // P), 3 * P * P
    outs = []
    for q8 in (0, 1):
        cols = torch.full((B * hp * wp, 2 * K), 7.0, dtype=torch.float16, device=DEV)
        check(lib().smk_im2col_f16(ptr(x), is_u8, ptr(cols), B, H, W, P, mean_std, q8, stream_ptr()), "smk_im2col_f16")
        torch.cuda.synchronize()
        outs.append(cols)
    xp = torch.nn.functional.pad(xn, (0, wp * P - W, 0, hp * P - H))
    ref = xp.view(B, 3, hp, P, wp, P).permute(0, 2, 4, 1, 3, 5).reshape(B * hp * wp, K)
    split, q = outs
    assert torch.equal(split[:, :K], ref.to(torch.float16))
    val = split[:, :K].float() + split[:, K:].float()                  # hi + lo
    assert (val - ref).abs().max().item() <= 1e-6
    hi, first, second = _unpack_q8(q, K)
    assert torch.equal(hi, split[:, :K].double()) and torch.equal(first, _q8(split[:, :K].float()))
    lo = split[:, K:].double()
    assert ((second * 2.0 ** -11 - lo).abs() <= lo.abs() * (2.0 ** -4 + 2.0 ** -9) + 2.0 ** -21).all()


@pytest.mark.parametrize("M_,N,K,epi,out_kind", [
    (300, 384, 384, 0, 1), (50432, 1536, 384, 1, 4), (50432, 384, 1536, 4, 1), (50176, 384, 768, 0, 1), (50432, 384, 384, 4, 1),
    (20000, 512, 1024, 0, 0), (130, 128, 64, 2, 3), (5000, 256, 192, 0, 4)])
def test_gemm_fp8_correction_terms(M_, N, K, epi, out_kind):
    """smk_gemm_q8: fp16 hi·hi + the two correction products on e4m3 operands (fp8 tensor-core rate, accumulated first at 2^15 and
    scaled into the fp16 term by scale-input-d).  Checked against (1) the exact value of the issued terms and (2) the fp64 product of
    the unsplit operands: ~fp32-grade like the 3-term fp16 split.  Shapes: fc1 (GELU, q8 output), fc2 (swap-AB, residual), patch
    embed, proj, a CTA-pair shape, small ragged ones."""
    torch.manual_seed(45)
    A32 = torch.randn(M_, K, device=DEV) * 1.5
    W32 = torch.randn(N, K, device=DEV) * 0.05
    bias = torch.randn(N, device=DEV)
    A2 = torch.zeros(M_, 2 * K, dtype=torch.float16, device=DEV)
    W2 = torch.zeros(N, 2 * K, dtype=torch.float16, device=DEV)
    check(lib().smk_split_q8(ptr(A32), K, ptr(A2), M_, K, 0, stream_ptr()), "smk_split_q8")
    check(lib().smk_split_q8(ptr(W32), K, ptr(W2), N, K, 1, stream_ptr()), "smk_split_q8")
    C0 = torch.randn(M_, N, device=DEV) if epi & 4 else None
    if out_kind == 1:
        out = C0.clone() if C0 is not None else torch.empty(M_, N, device=DEV)
    else:
        out = torch.full((M_, N * (1 if out_kind == 0 else 2)), 7.0, dtype=torch.float16, device=DEV)
    check(lib().smk_gemm_q8(ptr(A2), 2 * K, ptr(W2), 2 * K, ptr(bias), ptr(out), out.shape[1], M_, N, K, epi, out_kind, stream_ptr()), "smk_gemm_q8")
    torch.cuda.synchronize()
    ah, a1, a2 = _unpack_q8(A2, K)
    wh, w1, w2 = _unpack_q8(W2, K)
    issued = ah @ wh.t() + (a1 @ w1.t() + a2 @ w2.t()) * 2.0 ** -15 + bias.double()
    act = (lambda t: torch.nn.functional.gelu(t)) if epi & 1 else ((lambda t: torch.relu(t)) if epi & 2 else (lambda t: t))
    ref = act(issued) + (C0.double() if C0 is not None else 0)
    if out_kind == 1 or out_kind == 0:
        got = out.double()
    elif out_kind == 3:
        got = out[:, :N].double() + out[:, N:].double()
    else:
        hi, first, second = _unpack_q8(out, N)
        got = hi + second * 2.0 ** -11
        assert torch.equal(first, _q8(hi.float()))
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    tol = 2.0 ** -10 if out_kind == 0 else (2.0 ** -13 if out_kind == 4 else 2e-4)   # q8 output: lo keeps ~3 bits below the fp16 ulp
    assert err <= tol * scale, err
    # against the unsplit product: the dropped lo·lo term and the e4m3 rounding of the correction operands
    full = act(A32.double() @ W32.double().t() + bias.double()) + (C0.double() if C0 is not None else 0)
    err_full = (got - full).abs().max().item()
    assert err_full <= max(tol, 4e-4) * scale, err_full
    print(f"q8 gemm M={M_} N={N} K={K}: issued err {err / scale:.2e}, vs exact {err_full / scale:.2e}")


@pytest.mark.parametrize("B,N", [(3, 197), (2, 577)])
def test_attention_q8_output_equals_split_of_fp32_output(B, N):
    """Attention out mode 4 ([hi | e4m3 operands], the proj GEMM's A operand): identical to smk_split_q8 of the kernel's fp32-grade
    output ([hi | lo] mode 3 summed)."""
    torch.manual_seed(46)
    H, dh = 6, 64
    D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 1.5).to(torch.float16)
    o3 = torch.zeros(B * N, 2 * D, dtype=torch.float16, device=DEV)
    o4 = torch.zeros(B * N, 2 * D, dtype=torch.float16, device=DEV)
    for mode, o in ((3, o3), (4, o4)):
        if N <= 256:
            check(lib().smk_attention_tc_f16(ptr(qkv), ptr(o), 2 * D, mode, B, N, H, 0.125, stream_ptr()), "smk_attention_tc_f16")
        else:
            check(lib().smk_attention_tc_multi(ptr(qkv), 3 * D, ptr(qkv[:, D:]), 3 * D, ptr(qkv[:, 2 * D:]), 3 * D, B * N, B * N, N, N, 0, ptr(o), 2 * D,
                                               mode, B, N, N, H, 0.125, 1, stream_ptr()), "smk_attention_tc_multi")
    torch.cuda.synchronize()
    hi, first, second = _unpack_q8(o4, D)
    assert torch.equal(hi, o3[:, :D].double())
    assert torch.equal(first, _q8(o3[:, :D].float()))
    # lo of mode 3 is fp16(x − hi); mode 4 stores e4m3((x − hi)·2^11): equal up to the e4m3 rounding of the same residue
    lo = o3[:, D:].double()
    assert ((second * 2.0 ** -11 - lo).abs() <= lo.abs() * (2.0 ** -4 + 2.0 ** -9) + 2.0 ** -21).all()


@pytest.mark.parametrize("B,N,out_mode", [(3, 197, 0), (64, 197, 3), (2, 256, 3), (5, 130, 0), (150, 197, 3), (2, 17, 3)])
def test_attention_tcgen05_fp16_matches_torch(B, N, out_mode):
    """fp16 operands on the tcgen05 attention kernel (fp16s mode): plain fp16 output and the [hi | lo] fp16 split output."""
    torch.manual_seed(41)
    H, dh = 6, 64
    D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 1.5).to(torch.float16)
    ldo = D if out_mode == 0 else 2 * D
    out = torch.full((B * N, ldo), 7.0, device=DEV, dtype=torch.float16)
    check(lib().smk_attention_tc_f16(ptr(qkv), ptr(out), ldo, out_mode, B, N, H, 0.125, stream_ptr()), "smk_attention_tc_f16")
    torch.cuda.synchronize()
    q, k, v = [t.double().view(B, N, H, dh).transpose(1, 2) for t in qkv.view(B, N, 3 * D).split(D, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, D)
    got = out[:, :D].double() + (out[:, D:].double() if out_mode == 3 else 0)
    err = (got - ref).abs().max().item()
    # fp16 P (2^-11 relative) and ex2.approx; the plain output adds one fp16 rounding of values up to ~4
    assert err <= (6e-3 if out_mode == 0 else 4e-3), err
    assert (got - ref).abs().mean().item() <= 3e-4
    if out_mode == 3:
        assert (out[:, D:].float().abs() <= out[:, :D].float().abs() * 2.0 ** -10 + 1e-7).all()     # lo is the rounding residue of hi


@pytest.mark.parametrize("B,N,f16,out_mode", [(3, 577, 1, 3), (2, 785, 1, 3), (3, 577, 0, 0), (2, 785, 0, 0), (40, 577, 1, 3), (1, 176, 1, 0),
                                              (2, 177, 0, 1), (1, 352, 1, 3), (2, 600, 0, 2), (30, 785, 0, 0), (1, 1100, 1, 1)])
def test_attention_tcgen05_multi_tile_matches_torch(B, N, f16, out_mode):
    """Multi-key-tile tcgen05 attention (smk_attn_tc_multi.cu): 577 / 785-token encoders, online rescale of the TMEM accumulator across
    176-key tiles, shifted last tile (N = 177: 175 of its 176 columns are masked duplicates), exact multiples (176, 352), every output
    mode, both operand types; larger scores (x2) so that the running maximum really moves between tiles."""
    torch.manual_seed(43)
    H, dh = 6, 64
    D = H * dh
    dt = torch.float16 if f16 else torch.bfloat16
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 2.0).to(dt)
    parts = {0: 1, 1: 1, 2: 3, 3: 2}[out_mode]
    ldo = parts * D
    out = torch.full((B * N, ldo), 7.0, device=DEV, dtype=torch.float32 if out_mode == 1 else dt)
    check(lib().smk_attention_tc_multi(ptr(qkv), 3 * D, ptr(qkv[:, D:]), 3 * D, ptr(qkv[:, 2 * D:]), 3 * D, B * N, B * N, N, N, 0, ptr(out), ldo,
                                       out_mode, B, N, N, H, 0.125, f16, stream_ptr()), "smk_attention_tc_multi")
    torch.cuda.synchronize()
    q, k, v = [t.double().view(B, N, H, dh).transpose(1, 2) for t in qkv.view(B, N, 3 * D).split(D, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, D)
    if out_mode == 2:
        assert torch.equal(out[:, :D], out[:, D:2 * D])
        got = out[:, :D].double() + out[:, 2 * D:].double()
    elif out_mode == 3:
        got = out[:, :D].double() + out[:, D:].double()
    else:
        got = out.double()
    err = (got - ref).abs().max().item()
    tol = {(1, 0): 8e-3, (1, 1): 5e-3, (1, 3): 5e-3, (0, 0): 0.06, (0, 1): 0.04, (0, 2): 0.04}[(f16, out_mode)]
    assert err <= tol, err
    assert (got - ref).abs().mean().item() <= (4e-4 if f16 else 3e-3)


@pytest.mark.parametrize("nq,B,sub", [(20, 37, 0), (10, 5, 1), (32, 3, 0), (1, 2, 1), (20, 256, 1)])
def test_decoder_self_attention_fp32_matches_torch(nq, B, sub):
    """fp32 CUDA-core decoder self-attention; `sub`: V is the slice of a merged q|k|v projection (row stride 3D) and a per-query
    constant [nq, D] is subtracted from it on the way in."""
    torch.manual_seed(42)
    H, dh = 6, 64
    D = H * dh
    qkv = torch.randn(B * nq, 3 * D, device=DEV) * 3.0            # large scores: a peaked softmax, as with query_embed ~ N(0, 1)
    qkv[:, 2 * D:] /= 3.0
    vsub = torch.randn(nq, D, device=DEV) if sub else None
    out = torch.full((B * nq, 3 * D), 7.0, device=DEV, dtype=torch.bfloat16)
    check(lib().smk_dec_self_attention(ptr(qkv), 3 * D, ptr(qkv[:, 2 * D:]), 3 * D, ptr(vsub), ptr(out), B, nq, H, 0.125, stream_ptr()),
          "smk_dec_self_attention")
    torch.cuda.synchronize()
    qh = qkv[:, :D].double().view(B, nq, H, dh).transpose(1, 2)
    kh = qkv[:, D:2 * D].double().view(B, nq, H, dh).transpose(1, 2)
    v = qkv[:, 2 * D:].double().view(B, nq, D)
    if sub:
        v = v - vsub.double()[None]
    vh = v.view(B, nq, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B * nq, D)
    assert torch.equal(out[:, :D], out[:, D:2 * D])
    got = out[:, :D].double() + out[:, 2 * D:].double()
    assert (got - ref).abs().max().item() <= 1e-4           # fp32 math; the bf16 hi + lo pair keeps ~16 bits of values up to ~4


@pytest.mark.parametrize("Lq,Lk,kv_rows,kv_row0,q_f32", [(20, 196, 197, 1, 1), (20, 196, 197, 1, 0), (10, 144, 145, 1, 1), (32, 256, 256, 0, 1)])
def test_attention_small_fp16_matches_torch(Lq, Lk, kv_rows, kv_row0, q_f32):
    """fp16s-mode decoder cross-attention: fp16 K / V, the query either fp16 or fp32 rows rounded while staged; bf16 split output."""
    torch.manual_seed(43)
    B, H, dh = 37, 6, 64
    D = H * dh
    q32 = torch.randn(B * Lq, D, device=DEV)
    q16 = q32.to(torch.float16)
    kv = torch.randn(B * kv_rows, 2 * D, device=DEV).to(torch.float16)
    k_view, v_view = kv[:, :D], kv[:, D:]
    out = torch.full((B * Lq, 3 * D), 7.0, device=DEV, dtype=torch.bfloat16)
    qarg = q32 if q_f32 else q16
    check(lib().smk_attention_small_f16(ptr(qarg), D, C.c_void_p(k_view.data_ptr()), 2 * D, C.c_void_p(v_view.data_ptr()), 2 * D, kv_rows, kv_row0,
                                        ptr(out), 3 * D, 2, B, Lq, Lk, H, 0.125, q_f32, stream_ptr()))
    torch.cuda.synchronize()
    qh = q16.double().view(B, Lq, H, dh).transpose(1, 2)
    kk = kv.double().view(B, kv_rows, 2 * D)[:, kv_row0:kv_row0 + Lk]
    kh = kk[..., :D].reshape(B, Lk, H, dh).transpose(1, 2)
    vh = kk[..., D:].reshape(B, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(B * Lq, D)
    assert torch.equal(out[:, :D], out[:, D:2 * D])
    got = out[:, :D].double() + out[:, 2 * D:].double()
    assert (got - ref).abs().max().item() <= 3e-3            # fp16 P: 2^-11 relative; bf16 operands give ~2e-2 here


@pytest.mark.parametrize("B,Ra,a_row0,rows_a,hw", [(37, 120, 0, 120, 196), (5, 120, 100, 20, 196), (3, 60, 0, 60, 156), (2, 120, 0, 120, 576),
                                                   (150, 120, 0, 120, 196), (1, 384, 0, 384, 784)])
def test_gemm_batched_mask_logits_match_fp64(B, Ra, a_row0, rows_a, hw):
    """smk_gemm_batched — the mask-logit contraction as ONE batched tcgen05 GEMM: per image the (up to 128-row) tile of split query
    rows [hi | hi | lo] against that image's patch tokens [hi | lo] (cls row skipped), 3 terms, 3-D TMA store that clips pad rows."""
    torch.manual_seed(50)
    D, N = 384, hw + 1
    q = torch.randn(B * Ra, D, device=DEV) * 2.0
    t = torch.randn(B * N, D, device=DEV) * 1.5
    qh, ql = _split16(q, torch.bfloat16)
    th, tl = _split16(t, torch.bfloat16)
    q3 = torch.cat([qh, qh, ql], dim=1).contiguous()
    t2 = torch.cat([th, tl], dim=1).contiguous()
    out = torch.full((B, rows_a, hw), 7.0, device=DEV)
    ao, wo = (C.c_int32 * 3)(0, 0, 2 * D), (C.c_int32 * 3)(0, D, 0)
    check(lib().smk_gemm_batched(ptr(q3), 3 * D, B * Ra, Ra, a_row0, rows_a, ptr(t2), 2 * D, B * N, N, 1, hw, ptr(out), B, D, 0, 3, ao, wo, stream_ptr()),
          "smk_gemm_batched")
    torch.cuda.synchronize()
    qq = q.double().view(B, Ra, D)[:, a_row0:a_row0 + rows_a]
    tt = t.double().view(B, N, D)[:, 1:]
    ref = qq @ tt.transpose(1, 2)
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item() / 50), err          # 3-term bf16 split: ~2^-16 relative on |logit| up to ~150


def test_xattn_weight_folding_matches_fp64():
    """Wg = Wq_h^T Wk_h / 8, g = bq_h Wk_h / 8, Mcat = Wo_h Wv_h, bo2 = Wo bv + bo (smk_xattn_tc.cu): the algebra that removes the
    memory K/V projection from the decoder's cross-attention."""
    torch.manual_seed(60)
    D, H = 384, 6
    w_in = torch.randn(3 * D, D, device=DEV) * 0.03
    b_in = torch.randn(3 * D, device=DEV) * 0.02
    w_o, b_o = torch.randn(D, D, device=DEV) * 0.03, torch.randn(D, device=DEV) * 0.02
    wg = torch.empty(H * D, D, dtype=torch.float16, device=DEV)
    g, mcat, bo2 = torch.empty(H * D, device=DEV), torch.empty(D, H * D, device=DEV), torch.empty(D, device=DEV)
    check(lib().smk_xattn_fold_weights(ptr(w_in), ptr(b_in), ptr(w_o), ptr(b_o), ptr(wg), ptr(g), ptr(mcat), ptr(bo2), D, H, stream_ptr()))
    torch.cuda.synchronize()
    Wq, Wk, Wv = (t.double() for t in w_in.split(D))
    bq, bk, bv = (t.double() for t in b_in.split(D))
    for h in range(H):
        sl = slice(h * 64, (h + 1) * 64)
        G = Wq[sl].t() @ Wk[sl] / 8                     # [c', c]
        assert (wg[h * D:(h + 1) * D].double() - G.t()).abs().max().item() <= 2e-3 * G.abs().max().item() + 1e-7     # fp16 storage
        assert (g[h * D:(h + 1) * D].double() - bq[sl] @ Wk[sl] / 8).abs().max().item() <= 1e-7
        assert (mcat[:, h * D:(h + 1) * D].double() - w_o.double()[:, sl] @ Wv[sl]).abs().max().item() <= 1e-7
    assert (bo2.double() - (w_o.double() @ bv + b_o.double())).abs().max().item() <= 1e-7


@pytest.mark.parametrize("B,nq,hw", [(3, 20, 196), (150, 20, 196), (37, 10, 196), (5, 20, 156), (2, 21, 208), (1, 1, 16)])
def test_xattn_tcgen05_matches_torch(B, nq, hw):
    """Restructured decoder cross-attention on tcgen05: per image the nq·6 (query, head) rows against the image's patch tokens as
    keys and values (cls row skipped), fp16 operands, fp16 [hi | lo] split output."""
    torch.manual_seed(61)
    H, D, N = 6, 384, hw + 1
    qp = (torch.randn(B * nq, H * D, device=DEV) * 0.25).to(torch.float16)
    tok = (torch.randn(B * N, D, device=DEV) * 1.2).to(torch.float16)
    out = torch.full((B * nq, 2 * H * D), 7.0, device=DEV, dtype=torch.float16)
    check(lib().smk_xattn_tc(ptr(qp), ptr(tok), N, 1, ptr(out), B, nq, H, D, hw, stream_ptr()), "smk_xattn_tc")
    torch.cuda.synchronize()
    q = qp.double().view(B, nq * H, D)
    t = tok.double().view(B, N, D)[:, 1:]
    ref = (torch.softmax(q @ t.transpose(1, 2), -1) @ t).view(B * nq, H * D)
    HD = H * D
    got = out[:, :HD].double() + out[:, HD:].double()
    err = (got - ref).abs().max().item()
    assert err <= 4e-3, err                     # fp16 P (2^-11 relative), ex2.approx; values of magnitude ~1
    assert (got - ref).abs().mean().item() <= 2e-4
