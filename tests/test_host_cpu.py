"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the weight table matches the
reference state_dict schema, and the host-side finalisation reproduces the reference's metric values from
GPU-format records (emulated in numpy).  No compute call touches a GPU here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

import selfmask_b200 as S
from oracle import selfmask_oracle as O
from selfmask_b200 import metrics as M
from tests.helpers import s_measure_from_sums_loop
from tests.helpers import numpy_record

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(nq=20):
    return S.SmkConfig(patch=16, dim=384, depth=12, heads=6, mlp_dim=1536, n_queries=nq, dec_layers=6, dec_ffn=1536,
                       scale_factor=4, pos_grid=14)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "selfmask_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(smk_[a-z0-9_]+)\s*\(", header))
    from selfmask_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = S.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smk_version() >= 100


def test_weight_table_matches_reference_schema():
    for nq in (10, 20):
        cfg = O.make_config(n_queries=nq)
        schema = O.state_dict_schema(cfg)
        table = S.weight_table(_cfg(nq))
        assert [n for n, _, _ in table] == [n for n, _, _ in schema]
        assert len(table) == 267
        for (n, off, numel), (_, shape, _) in zip(table, schema):
            assert numel == int(np.prod(shape)), n
            assert off % 64 == 0
        offs = [o for _, o, _ in table]
        assert offs == sorted(offs)
        assert S.lib().smk_weights_numel(C.byref(_cfg(nq))) >= table[-1][1] + table[-1][2]


def test_bad_config_is_rejected_with_message():
    cfg = _cfg()
    cfg.dim = 100
    assert S.lib().smk_weight_count(C.byref(cfg)) < 0
    assert b"dim" in S.lib().smk_last_error()


def test_no_cpu_fallback():
    model = S.SelfMaskB200(n_queries=20)
    with pytest.raises(S.SmkError):
        model.to("cpu")
    with pytest.raises(S.SmkError):
        model(torch.zeros(1, 3, 224, 224))
    with pytest.raises(S.SmkError):
        S.compute_iou(torch.zeros(4, 4), torch.zeros(4, 4))


def test_finalize_reproduces_reference_metrics(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    recs = [numpy_record(p, gt) for p, gt in zip(g["preds"], g["gts"])]
    counts, sums = np.stack([r[0] for r in recs]), np.stack([r[1] for r in recs])
    f = M.finalize(counts, sums)
    # integer-derived metrics: bit-exact against the reference's own outputs
    assert np.array_equal(f["iou"], g["iou"])
    assert np.array_equal(f["f_score"], g["f_measure"])
    assert np.array_equal(f["f_max"], g["f_max"])
    assert np.array_equal(f["pixel_accuarcy"], g["pixel_acc"])
    fg_above = M.counts_above_thresholds(counts[:, 0:256].astype(np.int64))
    all_above = fg_above + M.counts_above_thresholds(counts[:, 256:512].astype(np.int64))
    assert np.array_equal(M.f_from_counts(fg_above, all_above, counts[:, 514][:, None].astype(np.int64)), g["fmax_vec"])
    # float reductions (different summation order than torch's float32 kernels): tolerance 2e-6 / 2e-5
    np.testing.assert_allclose(f["mae"], g["mae"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(f["f_mean"], g["f_mean"], rtol=0, atol=1e-6)
    a, b = f["s_measure"], g["s_measure"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=0, atol=2e-5)


def test_vectorised_s_measure_equals_scalar_definition(golden_dir):
    """The numpy-vectorised S-measure finalisation is bit-identical (NaNs included) to the per-record scalar version."""
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    recs = [numpy_record(p, gt) for p, gt in zip(g["preds"], g["gts"])]
    rng = np.random.default_rng(5)
    H, W = 40, 56
    for kind in range(12):                  # empty / full GT, single pixel, centroid on the border, constant predictions
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
        elif kind >= 9: gt = rng.random((H, W)) > 0.3 * (kind - 8)
        recs.append(numpy_record(p, gt))
    counts, sums = np.stack([r[0] for r in recs]), np.stack([r[1] for r in recs])
    a, b = M.s_measure_from_sums(counts, sums), s_measure_from_sums_loop(counts, sums)
    assert a.dtype == b.dtype == np.float64
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.isnan(a).any()
    assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    a2 = M.s_measure_from_sums(counts.reshape(-1, 2, 528)[: len(recs) // 2], sums.reshape(-1, 2, 32)[: len(recs) // 2])
    assert a2.shape == (len(recs) // 2, 2)


def test_histogram_counts_equal_direct_threshold_counts():
    rng = np.random.default_rng(0)
    thr = O.fmax_thresholds()
    p = rng.random((37, 53), dtype=np.float32)
    p.flat[:255] = thr                      # every threshold value itself must land in the right bin
    gt = rng.random((37, 53)) > 0.6
    counts, _ = numpy_record(p, gt)
    f = O.f_measure_all(p, gt)
    assert np.array_equal(M.counts_above_thresholds(counts[0:256].astype(np.int64)), f["tp_k"])
    assert np.array_equal(M.counts_above_thresholds((counts[0:256] + counts[256:512]).astype(np.int64)), f["tpfp_k"])


def test_running_mean_is_the_sequential_average_meter():
    rng = np.random.default_rng(1)
    vals = rng.random(1000).astype(np.float32)
    m = S.AverageMeter()
    for v in vals:
        m.update(np.asarray(v), 1)           # 0-d float32 arrays, like `tensor.numpy()` in the reference
    assert M.running_mean(vals) == float(m.avg)
    vals64 = rng.random(100)
    m = S.AverageMeter()
    for v in vals64:
        m.update(float(v), 1)
    assert M.running_mean(vals64) == float(m.avg)


def test_summarize_matches_reference_evaluator(golden_dir):
    """Host half of the evaluator: records built (in numpy) from the oracle's full-resolution masks must give
    the reference Evaluator's 14 averages."""
    ref = json.load(open(os.path.join(golden_dir, "evaluator.json")))
    cfg = O.make_config(n_queries=20)
    sd = O.synth_state_dict(cfg, seed=0)
    n, h, w = ref["n_img"], ref["h"], ref["w"]
    xs = O.normalize_images(O.synth_images_u8(n, h, w, seed=ref["image_seed"]))
    gts = O.synth_gt(n, h, w, seed=ref["gt_seed"], edge_every=ref["edge_every"])
    counts, sums = np.zeros((n, 2, 528), np.int32), np.zeros((n, 2, 32))
    with torch.no_grad():
        out = O.model_forward(sd, xs, cfg)
    full = O.upsample_bilinear(out["mask_pred"][:, -1].numpy(), 4)
    ob = out["objectness"][:, -1, :, 0].numpy()
    for i in range(n):
        inter, union = O.iou_counts(full[i], np.broadcast_to(gts[i, 0], full[i].shape))
        sel, ub = int(np.argmax(ob[i])), int(np.argmax(O.iou_from_counts(inter, union)))
        for j, q in enumerate((sel, ub)):
            counts[i, j], sums[i, j] = numpy_record(full[i, q], gts[i, 0])
    res = S.summarize(counts, sums)
    for k, v in ref["result"].items():
        assert abs(res[k] - v) <= 2e-5 * max(1.0, abs(v)), (k, res[k], v)


def test_shard_range_covers_everything():
    for n, w in [(5019, 8), (7, 8), (256, 1), (0, 4), (9, 2)]:
        spans = [S.shard_range(n, r, w) for r in range(w)]
        covered = [i for a, b in spans for i in range(a, b)]
        assert covered == list(range(n))


def test_package_synthetic_generators_match_the_oracle():
    from selfmask_b200 import synthetic as Y
    cfg = O.make_config(n_queries=20)
    a, b = O.synth_state_dict(cfg, seed=3), Y.synth_state_dict(S.weight_table(_cfg(20)), seed=3)
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape and torch.equal(a[k], b[k]) for k in a)
    assert np.array_equal(O.synth_images_u8(2, 64, 48, 5), Y.synth_images_u8(2, 64, 48, 5))
    assert np.array_equal(O.synth_gt(7, 64, 48, 6, edge_every=3), Y.synth_gt(7, 64, 48, 6, edge_every=3))


def test_x4_bilinear_cell_shortcut_premise():
    """smk_eval.cu's cell-classified IoU kernel skips cells whose 4 corner samples are all > 0.5 (or all <= 0.5).  That is exact
    iff the ATen-order bilinear blend f(a,b,c,d) = fma(top, ly0, bot*ly1), top = fma(a, lx0, b*lx1), is monotone in its corners
    (round-to-nearest is) and maps the constant field c to 0.5 for c = 0.5 and to > 0.5 for c = nextafter(0.5): checked here for
    every x4 weight pair, with float32 fma emulated exactly in float64."""
    f32 = np.float32

    def fma(a, b, c):          # a*b + c rounded once (a*b needs <= 27 bits, the f64 sum is exact)
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    def blend(c, lx1, ly1):
        lx0, ly0 = f32(1) - f32(lx1), f32(1) - f32(ly1)
        top = fma(c, lx0, f32(c * f32(lx1)))
        bot = top
        return fma(top, ly0, f32(bot * f32(ly1)))

    weights = [0.0, 0.125, 0.375, 0.625, 0.875]     # l1 of make_tap() at scale 4 (0 where the source index is clamped)
    half, above = f32(0.5), np.nextafter(f32(0.5), f32(1))
    rng = np.random.default_rng(0)
    for lx1 in weights:
        for ly1 in weights:
            assert blend(half, lx1, ly1) == half
            assert blend(above, lx1, ly1) > half
            for c in rng.random(64).astype(np.float32):          # constant fields never cross 0.5 from their own side
                assert (blend(c, lx1, ly1) > half) == (c > half)


def test_saliency_folder_reader_pairs_resizes_and_shards(tmp_path):
    """Directory reader for the batched evaluator: the reference's DUTS-TE layout, uint8 batches, {0,1} masks, contiguous shards."""
    from PIL import Image
    import selfmask_b200 as S
    d_img, d_gt = tmp_path / "DUTS-TE-Image", tmp_path / "DUTS-TE-Mask"
    d_img.mkdir()
    d_gt.mkdir()
    rng = np.random.default_rng(5)
    n = 7
    for i in range(n):
        h, w = int(rng.integers(40, 90)), int(rng.integers(40, 90))
        Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(d_img / f"im_{i:03d}.jpg")
        gt = np.zeros((h, w), np.uint8)
        gt[h // 4: h // 2 + i, w // 4: w // 2] = 255
        Image.fromarray(gt).save(d_gt / f"im_{i:03d}.png")
    ds = S.SaliencyFolder(str(tmp_path), "duts", img_size=32, batch_size=3, pin_memory=False)
    batches = list(ds)
    assert ds.n_images == n and len(ds) == 3 and [b["x"].shape[0] for b in batches] == [3, 3, 1]
    for b in batches:
        assert b["x"].dtype == torch.uint8 and b["x"].shape[1:] == (3, 32, 32)
        assert b["m"].dtype == torch.uint8 and b["m"].shape[1:] == (1, 32, 32) and set(np.unique(b["m"].numpy())) <= {0, 1}
        assert b["m"].sum() > 0
    assert [f for b in batches for f in b["filename"]] == [f"im_{i:03d}.jpg" for i in range(n)]
    # two ranks read disjoint contiguous shards that cover the set
    names = [[f for b in S.SaliencyFolder(str(tmp_path), "duts", 32, 2, rank=r, world_size=2, pin_memory=False) for f in b["filename"]] for r in range(2)]
    assert names[0] + names[1] == [f"im_{i:03d}.jpg" for i in range(n)] and len(names[0]) == 4
    with pytest.raises(S.SmkError):
        S.SaliencyFolder(str(tmp_path), "coco")


def _bf16_round(x):
    """float32 → bf16 (round to nearest even) → float32, in numpy."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_attention_online_softmax_scheme_matches_direct_softmax():
    """The encoder attention kernel (smk_attn_tc.cu) reads S once: every 16-key unit is exponentiated against the running
    row maximum rounded UP to an integer (log2 domain), kept as bf16, and corrected at the end by 2^(M_unit - M_final) — an
    exact power of two, so the corrected bf16 P equals bf16(exp2(x - M_final)) bit for bit; the row sum is the sum of the
    bf16 P (taken from the tensor core through a ones column).  Emulated here in numpy against the direct definition."""
    rng = np.random.default_rng(5)
    n_keys, scale_log2e = 197, 0.125 * 1.4426950408889634
    s = (rng.standard_normal((64, n_keys)) * rng.uniform(1, 60, (64, 1))).astype(np.float32)      # rows with very different spreads
    x = (s * np.float32(scale_log2e)).astype(np.float32)
    units = [np.arange(u, min(u + 16, n_keys)) for u in range(0, n_keys, 16)]
    M = np.full(64, -1e30, np.float32)
    p_units, m_units = [], []
    for cols in units:
        M = np.maximum(M, np.ceil(x[:, cols].max(1)))
        m_units.append(M.copy())
        p_units.append(_bf16_round(np.exp2((x[:, cols] - M[:, None]).astype(np.float32))))
    p = np.concatenate([_bf16_round(pu * np.exp2(mu - M)[:, None]) for pu, mu in zip(p_units, m_units)], axis=1)
    direct = _bf16_round(np.exp2((x - M[:, None]).astype(np.float32)))
    assert np.array_equal(p, direct)                      # the power-of-two correction adds no second rounding
    assert (M >= x.max(1)).all() and (M < x.max(1) + 1).all()
    probs = p / p.sum(1, keepdims=True)
    ref = np.exp(s.astype(np.float64) * 0.125)
    ref /= ref.sum(1, keepdims=True)
    assert np.abs(probs - ref).max() <= 4e-3              # bf16 P: half an ulp of values <= 1


def test_exp2_polynomial_of_the_attention_kernel():
    """ex2_poly (smk_attn_tc.cu; compiled out by default): round-to-nearest split through the 1.5*2^23 magic constant, degree-3
    minimax of 2^f on [-0.5, 0.5], exponent added with a shift — relative error below 1e-4 on the whole range the kernel clamps to."""
    x = np.linspace(-125, 0, 1_000_001).astype(np.float32)
    t = (x + np.float32(12582912.0)).astype(np.float32)
    f = (x - (t - np.float32(12582912.0))).astype(np.float32)
    p = np.float32(0.0551716648) * f + np.float32(0.2426111251)
    p = (p * f + np.float32(0.6932609677)).astype(np.float32)
    p = (p * f + np.float32(0.9999280572)).astype(np.float32)
    r = (p.view(np.int32) + (t.view(np.int32) << 23)).view(np.float32)
    assert np.abs(r / np.exp2(x.astype(np.float64)) - 1).max() <= 1e-4
