"""Host-side directory readers for the batched evaluator (BASELINE.json configs[4]).

Same directory layouts and file pairing as the reference's test-set readers — `datasets/duts.py:29-30`
(`DUTS-TE-Image/*.jpg`, `DUTS-TE-Mask/*.png`), `datasets/ecssd.py:17-18` (`images`, `ground_truth_mask`),
`datasets/dut_omron.py:17-18` (`DUT-OMRON-image`, `pixelwiseGT-new-PNG`) — sorted and paired by position.

Two protocols:
* `img_size=None` — the REFERENCE protocol (`datasets/duts.py:108-147`: test images at native resolution, batch 1, masks `> 0` →
  {0,1}): images keep their size; consecutive images of identical (H, W) are grouped into one batch (dataset order is preserved,
  so the ordered running means are the reference's), the model zero-pads to a multiple of the patch size
  (`vision_transformer.py:260-267`) and the evaluator crops `[..., :h, :w]` (evaluator.pyc@L209-211).  Metrics are comparable with
  the reference's.  Every distinct size is a model geometry: `SelfMaskB200(max_geometries=...)` bounds the workspaces kept.
* `img_size=S` — BASELINE.json configs[4] ("resized to 224x224"): every image AND mask is squashed to S x S without keeping the
  aspect ratio (bilinear; masks nearest) so that the whole sweep batches.  NOT the reference protocol: dataset metrics obtained
  this way are not comparable with published / native-resolution numbers (`SaliencyFolder.protocol` says which one ran).

Batches carry raw **uint8** pixels in page-locked memory: 1 byte per pixel-channel crosses PCIe and the loader's
`normalize(to_tensor(img))` (`datasets/base_dataset.py:250`) runs inside the patch im2col on the device, bit-identically.
Batches are the reference's dicts {'x', 'm', 'filename'}; multi-GPU runs read the contiguous shard `shard_range(n, rank, world)`.
"""
import os
from glob import glob
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from ._lib import SmkError
from .parallel import shard_range

LAYOUTS = {
    "duts": ("DUTS-TE-Image", "DUTS-TE-Mask"),
    "ecssd": ("images", "ground_truth_mask"),
    "dut_omron": ("DUT-OMRON-image", "pixelwiseGT-new-PNG"),
}


class SaliencyFolder:
    """Iterable of evaluation batches read from a DUTS-TE / ECSSD / DUT-OMRON style directory."""

    def __init__(self, dir_dataset: str, name: str = "duts", img_size: Optional[int] = 224, batch_size: int = 64, rank: int = 0,
                 world_size: int = 1, pin_memory: bool = True):
        if name not in LAYOUTS:
            raise SmkError(f"unknown dataset {name!r}; expected one of {sorted(LAYOUTS)}")
        d_img, d_gt = LAYOUTS[name]
        self.p_imgs: List[str] = sorted(glob(os.path.join(dir_dataset, d_img, "*.jpg")))
        self.p_gts: List[str] = sorted(glob(os.path.join(dir_dataset, d_gt, "*.png")))
        if len(self.p_imgs) != len(self.p_gts):
            raise SmkError(f"{len(self.p_imgs)} images but {len(self.p_gts)} masks under {dir_dataset}")
        self.name, self.img_size, self.batch_size = name, (int(img_size) if img_size else None), int(batch_size)
        self.protocol = "native resolution (reference protocol)" if self.img_size is None else f"squashed to {self.img_size}x{self.img_size} (not the reference protocol)"
        self.start, self.stop = shard_range(len(self.p_imgs), rank, world_size)
        self.pin_memory = pin_memory and torch.cuda.is_available()

    def __len__(self) -> int:                    # batches of this rank's shard (upper bound at native resolution)
        return -(-(self.stop - self.start) // self.batch_size)

    @property
    def n_images(self) -> int:
        return len(self.p_imgs)

    def _load(self, i: int) -> Tuple[np.ndarray, np.ndarray]:
        from PIL import Image
        s = self.img_size
        img, gt = Image.open(self.p_imgs[i]).convert("RGB"), Image.open(self.p_gts[i]).convert("L")
        if s is not None:
            img, gt = img.resize((s, s), Image.BILINEAR), gt.resize((s, s), Image.NEAREST)
        elif gt.size != img.size:
            raise SmkError(f"{self.p_gts[i]}: mask size {gt.size} differs from image size {img.size}")
        return np.asarray(img, np.uint8).transpose(2, 0, 1), (np.asarray(gt, np.uint8) > 0).astype(np.uint8)[None]

    def _emit(self, items: list) -> dict:
        n, (_, h, w) = len(items), items[0][1].shape
        x = torch.empty(n, 3, h, w, dtype=torch.uint8)
        m = torch.empty(n, 1, h, w, dtype=torch.uint8)
        if self.pin_memory:
            x, m = x.pin_memory(), m.pin_memory()
        for j, (_i, xi, mi) in enumerate(items):
            x[j] = torch.from_numpy(np.ascontiguousarray(xi))
            m[j] = torch.from_numpy(mi)
        return {"x": x, "m": m, "filename": [os.path.basename(self.p_imgs[i]) for i, _x, _m in items]}

    def __iter__(self) -> Iterator[dict]:
        if self.img_size is None:
            # native resolution: group CONSECUTIVE images of one size (order preserved), at most batch_size per batch
            group: list = []
            for i in range(self.start, self.stop):
                xi, mi = self._load(i)
                if group and (group[0][1].shape != xi.shape or len(group) >= self.batch_size):
                    yield self._emit(group)
                    group = []
                group.append((i, xi, mi))
            if group:
                yield self._emit(group)
            return
        s = self.img_size
        for b0 in range(self.start, self.stop, self.batch_size):
            idx = range(b0, min(b0 + self.batch_size, self.stop))
            x = torch.empty(len(idx), 3, s, s, dtype=torch.uint8)
            m = torch.empty(len(idx), 1, s, s, dtype=torch.uint8)
            if self.pin_memory:
                x, m = x.pin_memory(), m.pin_memory()
            for j, i in enumerate(idx):
                xi, mi = self._load(i)
                x[j] = torch.from_numpy(np.ascontiguousarray(xi))
                m[j] = torch.from_numpy(mi)
            yield {"x": x, "m": m, "filename": [os.path.basename(self.p_imgs[i]) for i in idx]}


def get_dataset(dir_dataset: str, dataset_name: str, img_size: Optional[int] = None, batch_size: int = 64, rank: int = 0,
                world_size: int = 1) -> SaliencyFolder:
    """Counterpart of the reference's dataset factory for the three test sets the evaluator sweeps (`utils/misc.py:43-151` with
    `eval_img_size=img_size`, evaluator.pyc@L176-180): `img_size=None` — the reference default — evaluates at native resolution."""
    return SaliencyFolder(dir_dataset, dataset_name, img_size, batch_size, rank, world_size)
