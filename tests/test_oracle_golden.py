"""Pin the CPU oracle (oracle/selfmask_oracle.py) against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import selfmask_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_thresholds_match_torch_arange(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    assert np.array_equal(O.fmax_thresholds(), g["thresholds"])
    assert len(O.fmax_thresholds()) == 255


def test_upsample_matches_aten_bit_exact(golden_dir):
    g = _load(golden_dir, "upsample.npz")
    # evaluator shapes (56²→224², 52×48→208×192): the nested-fma form is bit-exact with ATen
    assert np.array_equal(O.upsample_bilinear(g["probs"], 4), g["probs4"])
    assert np.array_equal(O.upsample_bilinear(g["odd"], 4), g["odd4"])
    # tiny planes take ATen's other rounding; both forms agree to 1 ulp
    assert np.array_equal(O.upsample_bilinear(g["x"], 4, variant="weights"), g["y4"])
    assert np.abs(O.upsample_bilinear(g["x"], 4) - g["y4"]).max() <= 1.2e-7


def test_metrics_match_reference(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    for i, (p, gt) in enumerate(zip(g["preds"], g["gts"])):
        m = O.image_metrics(p, gt)
        f = m["_f"]
        assert np.array_equal(m["iou"], g["iou"][i]), i
        assert np.array_equal(m["f_score"], g["f_measure"][i]), i
        assert np.array_equal(m["f_max"], g["f_max"][i]), i
        assert np.array_equal(m["f_mean"], g["f_mean"][i]), i
        assert np.array_equal(O.f_from_counts(f["tp_k"], f["tpfp_k"], f["tp_fn"]), g["fmax_vec"][i]), i
        assert np.array_equal(m["mae"], g["mae"][i]), i
        assert np.array_equal(m["pixel_accuarcy"], g["pixel_acc"][i]), i
        a, b = m["s_measure"], float(g["s_measure"][i])
        assert (np.isnan(a) and np.isnan(b)) or a == b, (i, a, b)


@pytest.mark.parametrize("name,nq,b,h,w", [("nq20_224", 20, 2, 224, 224), ("nq10_224", 10, 1, 224, 224),
                                           ("nq20_384", 20, 1, 384, 384), ("nq20_200x180", 20, 1, 200, 180)])
def test_model_matches_reference(golden_dir, name, nq, b, h, w):
    g = _load(golden_dir, f"model_{name}.npz")
    cfg = O.make_config(n_queries=nq)
    sd = O.synth_state_dict(cfg, seed=0)
    x = O.normalize_images(O.synth_images_u8(b, h, w, seed=1234))
    with torch.no_grad():
        out = O.model_forward(sd, x, cfg)
    mp = out["mask_pred"].numpy()
    assert mp.shape[1:3] == (6, nq)
    # same math, same ATen kernels, different association in a few places → fp32 noise only
    np.testing.assert_allclose(mp[:, -1], g["mask_pred_last"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(mp[:, :, :, ::7, ::5], g["mask_pred_sub"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(mp.mean(axis=(-1, -2)), g["mask_pred_layer_means"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out["objectness"].numpy(), g["objectness"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(out["features"].numpy(), g["features"], atol=2e-5, rtol=0)
    assert np.array_equal(out["objectness"].numpy()[:, -1, :, 0].argmax(-1), g["objectness"][:, -1, :, 0].argmax(-1))


def test_evaluator_matches_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "evaluator.json")))
    cfg = O.make_config(n_queries=20)
    sd = O.synth_state_dict(cfg, seed=0)
    n, h, w = ref["n_img"], ref["h"], ref["w"]
    xs = O.normalize_images(O.synth_images_u8(n, h, w, seed=ref["image_seed"]))
    gts = O.synth_gt(n, h, w, seed=ref["gt_seed"], edge_every=ref["edge_every"])
    with torch.no_grad():
        res = O.evaluate(lambda x: O.model_forward(sd, x, cfg), [(xs[i:i + 1], gts[i:i + 1]) for i in range(n)])
    for k, v in ref["result"].items():
        assert abs(res[k] - v) <= 2e-5 * max(1.0, abs(v)), (k, res[k], v)
    # the integer side must agree exactly: same selected / upper-bound queries ⇒ same count-derived values
    assert len(res["_images"]) == n
