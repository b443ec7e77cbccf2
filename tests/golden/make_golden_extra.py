"""Round-2 additions to the golden fixtures, generated like make_golden.py (the UNMODIFIED reference, imported read-only from
/root/reference with the same shims) without touching the round-1 files:

  model_vits8_sf2_224.npz   the configuration the reference ships (configs/*.yaml:14,39: ViT-S/8, scale_factor 2) — 785 tokens
  encoder_only.npz / .json  what MaskFormer.forward(x, encoder_only=True) (maskformer.py:183-189) returns — or how it fails

    python tests/golden/make_golden_extra.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import bootstrap_reference, ref_model  # noqa: E402
from oracle import selfmask_oracle as O  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    misc = bootstrap_reference()
    # ---- ViT-S/8, scale factor 2 (the shipped yaml geometry) ---------------------------------------------------------------
    cfg = O.make_config(n_queries=20, patch_size=8, scale_factor=2, pos_grid=28)
    sd = O.synth_state_dict(cfg, seed=3)
    model = ref_model(misc, cfg, sd)
    x = O.normalize_images(O.synth_images_u8(1, 224, 224, seed=55))
    with torch.no_grad():
        out = model(x)
    mp, ob = out["mask_pred"], out["objectness"]
    np.savez_compressed(f"{HERE}/model_vits8_sf2_224.npz", mask_pred_last=mp[:, -1].numpy(), mask_pred_sub=mp[:, :, :, ::7, ::5].numpy(),
                        mask_pred_layer_means=mp.mean(dim=(-1, -2)).numpy(), objectness=ob.numpy(), features=out["features"].numpy())
    print("vits8", tuple(mp.shape), tuple(ob.shape))
    # ---- encoder_only ------------------------------------------------------------------------------------------------------
    cfg16 = O.make_config(n_queries=20)
    sd16 = O.synth_state_dict(cfg16, seed=0)
    m16 = ref_model(misc, cfg16, sd16)
    x16 = O.normalize_images(O.synth_images_u8(2, 224, 224, seed=1234))
    info = {"call": "MaskFormer.forward(x, encoder_only=True)  (networks/maskformer/maskformer.py:183-189)"}
    try:
        with torch.no_grad():
            eo = m16(x16, encoder_only=True)
        pt = eo["patch_tokens"]
        info.update(status="ok", shape=list(pt.shape), keys=sorted(eo.keys()))
        np.savez_compressed(f"{HERE}/encoder_only.npz", patch_tokens=pt.numpy())
    except Exception as e:   # the reference `.view`s a non-contiguous b x D x hw tensor as b x h x w x D
        info.update(status="raises", error_type=type(e).__name__, error=str(e)[:300])
        with torch.no_grad():
            feats = m16.forward_encoder(x16)[:, -1]            # b x D x hw: what the failing line holds
        np.savez_compressed(f"{HERE}/encoder_only.npz", last_layer_features=feats.numpy())
    json.dump(info, open(f"{HERE}/encoder_only.json", "w"), indent=1)
    print(json.dumps(info, indent=1))


if __name__ == "__main__":
    main()
