"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run once in the build container (the reference is not available on the GPU box):

    python tests/golden/make_golden.py

It imports ``/root/reference`` read-only with the import shims of SURVEY.md §8c / Appendix A
(stubs for packages that are not installed and are not on the hot path, ``load_model`` → no-op so
that no download is attempted, ``SMeasure.cuda`` follows ``torch.cuda.is_available()``, sourceless
import of ``evaluator.pyc`` + the matching ``base_structure.pyc``), loads the oracle's synthetic
weights into the reference ``MaskFormer`` with ``load_state_dict(strict=True)`` and stores what the
reference computes.  Nothing from the reference is copied into the repo; only its *outputs* are.
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import selfmask_oracle as O  # noqa: E402

REF = "/root/reference"


def bootstrap_reference():
    ora = tempfile.mkdtemp(prefix="smk_ref_")
    shutil.copy(f"{REF}/__pycache__/evaluator.cpython-312.pyc", f"{ora}/evaluator.pyc")
    shutil.copy(f"{REF}/__pycache__/base_structure.cpython-312.pyc", f"{ora}/base_structure.pyc")
    sys.path[:0] = [ora, REF]
    for name, attrs in [("natsort", {"natsorted": sorted}), ("networks.timm_deit", {}), ("matplotlib", {}),
                        ("matplotlib.pyplot", {}), ("pycocotools", {}), ("pycocotools.mask", {"decode": None})]:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.modules["ujson"] = json
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import utils.misc as misc
    misc.load_model = lambda *a, **k: None
    import metrics.s_measure as sm
    orig = sm.SMeasure.__init__

    def init(self, alpha=0.5):
        orig(self, alpha)
        self.cuda = torch.cuda.is_available()
    sm.SMeasure.__init__ = init
    return misc


def ref_model(misc, cfg, sd):
    import yaml
    from argparse import Namespace
    y = yaml.safe_load(open(f"{REF}/configs/duts-dino-k234-nq20-224-swav-mocov2-dino-p16-sr10100.yaml"))
    y.update(patch_size=cfg["patch_size"], scale_factor=cfg["scale_factor"], n_queries=cfg["n_queries"])
    model = misc.get_model(arch="maskformer", configs=Namespace(**y)).eval()
    model.load_state_dict(sd, strict=True)
    return model


def metric_cases():
    """Probability maps that stress the metric code: smooth blobs, values exactly on F-max
    thresholds and on 0.5, saturated 0/1 pixels, empty / full GT, a GT whose centroid rounds
    half-to-even."""
    rng = np.random.default_rng(7)
    H, W = 64, 80
    thr = O.fmax_thresholds()
    preds, gts = [], []
    for i in range(10):
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        p = 1 / (1 + np.exp(-(rng.normal(0, 3) + 6 * np.exp(-((yy - rng.uniform(10, 50)) ** 2 + (xx - rng.uniform(10, 70)) ** 2)
                                                          / (2 * rng.uniform(5, 20) ** 2)) + rng.normal(0, 1.5, (H, W)))))
        p = p.astype(np.float32)
        idx = rng.integers(0, H * W, 400)
        p.flat[idx[:200]] = thr[rng.integers(0, 255, 200)]      # exactly on thresholds
        p.flat[idx[200:260]] = 0.5
        p.flat[idx[260:330]] = 0.0
        p.flat[idx[330:]] = 1.0
        g = O.synth_gt(1, H, W, seed=100 + i)[0, 0]
        if i == 7:
            g[:] = 0
        if i == 8:
            g[:] = 1
        if i == 9:                                               # centroid exactly x.5 → half-even
            g[:] = 0
            g[10:13, 20:22] = 1
        preds.append(p)
        gts.append(g)
    return np.stack(preds), np.stack(gts)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    misc = bootstrap_reference()
    from metrics.iou import compute_iou
    from metrics.f_measure import FMeasure
    from metrics.mae import compute_mae
    from metrics.pixel_acc import compute_pixel_accuracy
    from metrics.s_measure import SMeasure

    # ---- metrics ------------------------------------------------------------------------
    preds, gts = metric_cases()
    rec = {k: [] for k in ("iou", "f_measure", "f_max", "f_mean", "mae", "pixel_acc", "s_measure", "fmax_vec")}
    for p, g in zip(preds, gts):
        pt, gt64 = torch.from_numpy(p), torch.from_numpy(g.astype(np.int64))
        rec["iou"].append(compute_iou(pt, gt64).numpy())
        fm = FMeasure()
        f = fm(pt, gt64)
        for k in ("f_measure", "f_max", "f_mean"):
            rec[k].append(f[k].numpy())
        thr = torch.arange(0, 1, 1 / 255).view(255, 1, 1)
        rec["fmax_vec"].append(fm._compute_f_measure(pt[None].repeat(255, 1, 1), gt64[None].repeat(255, 1, 1), thr).numpy())
        rec["mae"].append(compute_mae(pt, gt64).numpy())
        rec["pixel_acc"].append(compute_pixel_accuracy(pt, gt64).numpy())
        rec["s_measure"].append(SMeasure()(pred_mask=pt.clone(), gt_mask=gt64.to(torch.float32)))
    np.savez_compressed(f"{HERE}/metrics.npz", preds=preds, gts=gts,
                        thresholds=torch.arange(0, 1, 1 / 255).numpy(),
                        **{k: np.asarray(v) for k, v in rec.items()})
    # bilinear ×4 of probabilities, as the evaluator does it (evaluator.pyc@L209-211)
    interp = torch.nn.functional.interpolate
    r3 = np.random.default_rng(3)
    small = torch.from_numpy(r3.random((2, 3, 9, 11), dtype=np.float32))
    probs = torch.sigmoid(torch.from_numpy((r3.random((3, 56, 56), dtype=np.float32) - 0.5) * 40))[None]
    odd = torch.sigmoid(torch.from_numpy((r3.random((2, 52, 48), dtype=np.float32) - 0.5) * 40))[None]
    np.savez_compressed(f"{HERE}/upsample.npz", x=small.numpy(),
                        y4=interp(small, scale_factor=4, mode="bilinear", align_corners=False).numpy(),
                        probs=probs.numpy(), probs4=interp(probs, scale_factor=4, mode="bilinear", align_corners=False).numpy(),
                        odd=odd.numpy(), odd4=interp(odd, scale_factor=4, mode="bilinear", align_corners=False).numpy())

    # ---- model --------------------------------------------------------------------------
    cases = [("nq20_224", dict(n_queries=20), 2, 224, 224),
             ("nq10_224", dict(n_queries=10), 1, 224, 224),
             ("nq20_384", dict(n_queries=20), 1, 384, 384),
             ("nq20_200x180", dict(n_queries=20), 1, 200, 180)]
    models = {}
    for name, kw, b, h, w in cases:
        cfg = O.make_config(**kw)
        sd = O.synth_state_dict(cfg, seed=0)
        model = models.get(kw["n_queries"]) or ref_model(misc, cfg, sd)
        models[kw["n_queries"]] = model
        x = O.normalize_images(O.synth_images_u8(b, h, w, seed=1234))
        with torch.no_grad():
            out = model(x)
        mp, ob = out["mask_pred"], out["objectness"]
        np.savez_compressed(f"{HERE}/model_{name}.npz",
                            mask_pred_last=mp[:, -1].numpy(),
                            mask_pred_layer_means=mp.mean(dim=(-1, -2)).numpy(),
                            mask_pred_sub=mp[:, :, :, ::7, ::5].numpy(),
                            objectness=ob.numpy(), features=out["features"].numpy())
        print(name, tuple(mp.shape), tuple(ob.shape), "logit-ish range", float(mp.min()), float(mp.max()))

    # ---- evaluator (reference Evaluator.__call__ loop, dataset layer replaced by tensors) --
    import evaluator as ref_eval
    cfg = O.make_config(n_queries=20)
    n_img, h, w = 6, 224, 224
    xs = O.normalize_images(O.synth_images_u8(n_img, h, w, seed=77))
    gts_e = O.synth_gt(n_img, h, w, seed=78, edge_every=3)        # idx 2 empty, idx 5 full

    class FakeDataset:
        def get_dataloader(self, **kw):
            def it():
                for i in range(n_img):
                    yield {"x": xs[i:i + 1], "m": torch.from_numpy(gts_e[i:i + 1].astype(np.int64))}
            return it(), range(n_img)

    class Bar(list):
        def set_description(self, *_a, **_k):
            pass
    FakeDataset.get_dataloader = (lambda self, **kw: (iter([{"x": xs[i:i + 1], "m": torch.from_numpy(gts_e[i:i + 1].astype(np.int64))}
                                                            for i in range(n_img)]), Bar(range(n_img))))
    ref_eval.get_dataset = lambda **kw: FakeDataset()
    out_dir = tempfile.mkdtemp(prefix="smk_eval_")
    ev = ref_eval.Evaluator(network=models[20], dir_dataset=out_dir, visualizer=None)
    ev._visualize = lambda *a, **k: None
    res = ev(dataset_name="duts", dir_ckpt=out_dir, batch_size=1, device=torch.device("cpu"))
    txt = open(f"{out_dir}/metrics_duts.txt").read()
    json.dump({"result": {k: float(v) for k, v in res.items()}, "metrics_txt": txt, "n_img": n_img, "h": h, "w": w,
               "image_seed": 77, "gt_seed": 78, "edge_every": 3},
              open(f"{HERE}/evaluator.json", "w"), indent=1)
    print(json.dumps({k: float(v) for k, v in res.items()}, indent=1))
    print(txt)


if __name__ == "__main__":
    main()
