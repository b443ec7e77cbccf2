"""Synthetic weights / images / ground truth of BASELINE.json's shapes (there are no datasets or checkpoints
offline).  Deterministic numpy PCG64 streams; the oracle regenerates the same values independently
(tests/test_host_cpu.py checks that both generators agree)."""
from typing import Dict, List, Tuple

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_SCALE = {"w": 0.02, "w2": 0.03, "b": 0.02, "ln_b": 0.02, "emb": 1.0}


def _kind(name: str) -> str:
    if name == "query_embed":
        return "emb"
    if ".norm" in name or name.startswith(("encoder.norm", "decoder.norm")):
        return "ln_w" if name.endswith("weight") else "ln_b"
    if name.endswith("bias"):
        return "b"
    if name.startswith("encoder.") or name.endswith("linear2.weight"):
        return "w"
    return "w2"


def _shape(name: str, numel: int, dim: int, patch: int) -> Tuple[int, ...]:
    if name == "encoder.cls_token":
        return (1, 1, dim)
    if name == "encoder.pos_embed":
        return (1, numel // dim, dim)
    if name == "encoder.patch_embed.proj.weight":
        return (dim, 3, patch, patch)
    if name.endswith("bias") or ".norm" in name or "norm." in name:
        return (numel,)
    if name == "ffn.layers.2.weight":
        return (1, dim)
    if name.endswith(("fc2.weight", "linear2.weight")):
        return (dim, numel // dim)
    return (numel // dim, dim)


def synth_state_dict(table: List[Tuple[str, int, int]], dim: int = 384, patch: int = 16, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights under the reference's state_dict names, in weight-table order."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, _off, numel in table:
        shape = _shape(name, numel, dim, patch)
        k = _kind(name)
        a = (1.0 + 0.05 * rng.standard_normal(shape, dtype=np.float32)) if k == "ln_w" else _SCALE[k] * rng.standard_normal(shape, dtype=np.float32)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    return sd


def synth_images_u8(n: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, 3, h, w), np.uint8)
    for i in range(n):
        img = rng.uniform(0, 255, (3, 1, 1)).astype(np.float32) * np.ones((3, h, w), np.float32)
        for _ in range(4):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            s = rng.uniform(0.08, 0.35) * max(h, w)
            amp = rng.uniform(-160, 160, (3, 1, 1)).astype(np.float32)
            img += amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))[None]
        img += rng.normal(0, 12, (3, h, w)).astype(np.float32)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def normalize_images(u8: np.ndarray) -> torch.Tensor:
    x = torch.from_numpy(u8.astype(np.float32)) / 255.0
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def synth_gt(n: int, h: int, w: int, seed: int = 4321, edge_every: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((n, 1, h, w), np.uint8)
    for i in range(n):
        y0, x0 = int(rng.integers(h // 8, h // 2)), int(rng.integers(w // 8, w // 2))
        rh, rw = int(rng.integers(h // 8, h // 2)), int(rng.integers(w // 8, w // 2))
        m = np.zeros((h, w), bool)
        m[y0:y0 + rh, x0:x0 + rw] = True
        cy, cx = rng.uniform(h * 0.3, h * 0.7), rng.uniform(w * 0.3, w * 0.7)
        ay, ax = rng.uniform(h * 0.08, h * 0.3), rng.uniform(w * 0.08, w * 0.3)
        m |= ((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2 <= 1.0
        if edge_every and (i % (2 * edge_every)) == edge_every - 1:
            m[:] = False
        elif edge_every and (i % (2 * edge_every)) == 2 * edge_every - 1:
            m[:] = True
        out[i, 0] = m
    return out
