"""Host-side mirror of the reference model interface for the B200 path.

`SelfMaskB200` keeps the constructor keywords, attributes and `forward` contract of
`networks/maskformer/maskformer.py:11-72,164-251` (MaskFormer) so it can stand in for it behind
`BaseStructure._forward` (base_structure.py:18-24) and `Evaluator(network=...)`:

    model(x, encoder_only=False, skip_decoder=False) -> {"mask_pred", "objectness", "features"}

All arithmetic happens in libselfmask_b200.so (hand-written sm_100a CUDA); torch only owns device memory
and the stream.
"""
import collections
import ctypes as C
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import SmkConfig, SmkError, check, lib, ptr, stream_ptr

# "fp16s" (default): fp16 tensor-core operands with per-contraction split terms — the mode that meets north_star's 2e-2 logit /
# 99.9 % IoU-agreement criteria at speed; "bf16": single-pass bf16 (fastest, outside that tolerance); "bf16x3": every GEMM a 3-term
# bf16 split; "fp32": CUDA-core validation mode
_MODES = {"fp32": _lib.SMK_MODE_FP32, "bf16": _lib.SMK_MODE_BF16, "bf16x3": _lib.SMK_MODE_BF16X3, "fp16s": _lib.SMK_MODE_FP16S}


def weight_table(cfg: SmkConfig):
    """[(state_dict key, offset, numel)] of the canonical fp32 weight blob (from the C library)."""
    n = lib().smk_weight_count(C.byref(cfg))
    if n <= 0:
        check(n, "smk_weight_count")
    out, buf = [], C.create_string_buffer(256)
    off, num = C.c_int64(), C.c_int64()
    for i in range(n):
        check(lib().smk_weight_entry(C.byref(cfg), i, buf, 256, C.byref(off), C.byref(num)), "smk_weight_entry")
        out.append((buf.value.decode(), off.value, num.value))
    return out


class SelfMaskB200(torch.nn.Module):
    """Drop-in for the reference `MaskFormer` (vit_small / dino, use_binary_classifier=True).

    Extra keywords (not in the reference): `mode` ("fp16s" parity at speed, the default | "bf16" fastest, outside the logit
    tolerance | "bf16x3" | "fp32" validation),
    `return_intermediate` False gives the legal fast variant of the interface — 4-D `mask_pred` with 3-D
    `objectness` (evaluator.pyc@L199-205), `max_batch` sizes the workspace.
    """

    def __init__(self, n_queries: int = 20, arch: str = "vit_small", patch_size: int = 16, training_method: str = "dino",
                 n_decoder_layers: int = 6, normalize_before: bool = False, return_intermediate: bool = True,
                 learnable_pixel_decoder: bool = False, lateral_connection: bool = False, scale_factor: int = 4,
                 abs_2d_pe_init: bool = False, use_binary_classifier: bool = True, mode: str = "fp16s", max_batch: int = 64,
                 device: Optional[torch.device] = None, max_geometries: int = 4):
        super().__init__()
        if arch != "vit_small" or training_method != "dino":
            raise SmkError("only arch='vit_small', training_method='dino' is on the B200 path (SURVEY.md §8)")
        if normalize_before or lateral_connection or learnable_pixel_decoder:
            raise SmkError("normalize_before / lateral_connection / learnable_pixel_decoder are not on the path")
        if not use_binary_classifier:
            raise SmkError("use_binary_classifier=False is not on the path (the evaluator requires objectness)")
        if mode not in _MODES:
            raise SmkError(f"mode must be one of {sorted(_MODES)}")
        self.arch, self.use_binary_classifier = arch, True
        self.lateral_connection, self.learnable_pixel_decoder = False, False
        self.scale_factor, self.return_intermediate, self.mode = scale_factor, return_intermediate, mode
        self.cfg = SmkConfig(patch=patch_size, dim=384, depth=12, heads=6, mlp_dim=1536, n_queries=n_queries,
                             dec_layers=n_decoder_layers, dec_ffn=1536, scale_factor=scale_factor,
                             pos_grid=224 // patch_size)
        # attributes the reference callers read (maskformer.py:32-34,107,184-185)
        self.encoder = SimpleNamespace(patch_size=patch_size, depth=12, n_embs=384, n_heads=6, mlp_ratio=4)
        self.max_batch = max_batch
        # uint8 inputs: the loader's normalisation constants (datasets/duts.py ImageNet mean / std, applied at base_dataset.py:250)
        self.pixel_mean, self.pixel_std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
        self._table = None
        self._blob = None          # fp32 weights, one device tensor in the library's canonical order
        # (H, W) -> (handle, workspace tensor, capacity), least recently used first.  A workspace is sized for ONE geometry
        # (1.3 GB at 224^2 x 256 images), and the reference protocol evaluates at native resolution — hundreds of distinct sizes on
        # DUTS-TE / DUT-OMRON — so the cache is bounded: beyond `max_geometries` the least recently used workspace is released,
        # and a new geometry's capacity is the batch actually seen (not `max_batch`) once more than one geometry is live.
        self._handles = collections.OrderedDict()
        self.max_geometries = max(1, int(max_geometries))
        self._graphs = {}          # predict_one: (H, W, dtype) -> (graph, static input, static outputs)
        self._device = torch.device(device) if device is not None else None
        self._loaded = False

    # ---- nn.Module protocol used by the reference (evaluator.pyc@L358-360, app.py:178-187) ----------
    def to(self, device=None, *a, **k):
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise SmkError("SelfMaskB200 only runs on CUDA devices")
            if self._blob is not None and self._blob.device != device:
                self._release()
                self._blob = self._blob.to(device)
            self._device = device
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def table(self):
        if self._table is None:
            self._table = weight_table(self.cfg)
        return self._table

    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts the reference's raw state_dict or a checkpoint dict with a 'model' entry."""
        if "model" in state_dict and "query_embed" not in state_dict:
            state_dict = state_dict["model"]
        dev = self._device or torch.device("cuda", torch.cuda.current_device())
        table = self.table()
        names = {n for n, _, _ in table}
        missing = sorted(names - set(state_dict.keys()))
        unexpected = sorted(set(state_dict.keys()) - names)
        if strict and (missing or unexpected):
            raise SmkError(f"load_state_dict: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
        total = lib().smk_weights_numel(C.byref(self.cfg))
        host = torch.zeros(total, dtype=torch.float32)
        for name, off, numel in table:
            if name not in state_dict:
                continue
            t = state_dict[name].detach().to("cpu", torch.float32).reshape(-1)
            if t.numel() != numel:
                raise SmkError(f"load_state_dict: {name} has {t.numel()} elements, expected {numel}")
            host[off:off + numel] = t
        self._release()
        self._blob = host.to(dev)
        self._device = dev
        self._loaded = True
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def state_dict(self, *a, **k):
        if self._blob is None:
            return {}
        return {n: self._blob[o:o + m].clone() for n, o, m in self.table()}

    def _release(self):
        self._graphs = {}
        for ent in self._handles.values():
            lib().smk_model_destroy(ent[0])
        self._handles = collections.OrderedDict()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _handle(self, B: int, H: int, W: int):
        if not self._loaded:
            raise SmkError("weights not loaded: call load_state_dict first")
        key = (H, W)
        ent = self._handles.get(key)
        if ent is not None and ent[2] >= B:
            self._handles.move_to_end(key)
            return ent[0]
        # a handle is about to be destroyed / replaced: graphs captured on it replay freed memory, and earlier launches must be done
        if ent is not None or len(self._handles) >= self.max_geometries:
            torch.cuda.current_stream(self._blob.device).synchronize()
        if ent is not None:
            lib().smk_model_destroy(ent[0])
            del self._handles[key]
            self._graphs = {k: v for k, v in self._graphs.items() if k[:2] != key}
        while len(self._handles) >= self.max_geometries:
            old_key, old = self._handles.popitem(last=False)
            lib().smk_model_destroy(old[0])
            self._graphs = {k: v for k, v in self._graphs.items() if k[:2] != old_key}
        cap = max(B, self.max_batch) if not self._handles else B
        mode = _MODES[self.mode]
        nbytes = lib().smk_model_workspace_bytes(C.byref(self.cfg), mode, cap, H, W)
        if nbytes <= 0:
            check(int(nbytes), "smk_model_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self._blob.device)
        handle = C.c_void_p()
        check(lib().smk_model_create(C.byref(self.cfg), mode, ptr(self._blob), ptr(ws), nbytes, cap, H, W, stream_ptr(),
                                     C.byref(handle)), "smk_model_create")
        self._handles[key] = (handle, ws, cap)
        return handle

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, encoder_only: bool = False, skip_decoder: bool = False) -> Dict[str, torch.Tensor]:
        """x: b x 3 x H x W float32 on the GPU (maskformer.py:164), ImageNet-normalised as the reference's loader
        leaves it (datasets/base_dataset.py:250) — or uint8 raw pixels, in which case that normalisation
        (`self.pixel_mean` / `self.pixel_std`) is fused into the patch-embedding im2col, bit-identically.
        `skip_decoder` is accepted and ignored, as in the reference (:118-135)."""
        if x.ndim != 4 or x.shape[1] != 3:
            raise SmkError(f"expected b x 3 x H x W, got {tuple(x.shape)}")
        _lib.require_cuda(x, "x")
        is_u8 = x.dtype == torch.uint8
        x = x.contiguous() if is_u8 else x.contiguous().float()
        B, _, H, W = x.shape
        cfg = self.cfg
        hp, wp = -(-H // cfg.patch), -(-W // cfg.patch)
        with torch.cuda.device(x.device):
            handle = self._handle(B, H, W)
            L = cfg.dec_layers if (self.return_intermediate and not encoder_only) else 1
            ho, wo = hp * cfg.scale_factor, wp * cfg.scale_factor
            def run(all_layers, mp, ob, ft):
                if is_u8:
                    ms = (C.c_float * 6)(*self.pixel_mean, *self.pixel_std)
                    check(lib().smk_model_forward_u8(handle, ptr(x), ms, B, H, W, all_layers, mp, ob, ft, stream_ptr()),
                          "smk_model_forward_u8")
                else:
                    check(lib().smk_model_forward(handle, ptr(x), B, H, W, all_layers, mp, ob, ft, stream_ptr()), "smk_model_forward")
            if encoder_only:
                run(0, None, None, None)
                tok = torch.empty(B, hp * wp + 1, cfg.dim, dtype=torch.float32, device=x.device)
                check(lib().smk_model_tap(handle, 1, ptr(tok), tok.numel(), stream_ptr()), "smk_model_tap")
                # maskformer.py:183-189 (the reference `.view`s a b x D x hw tensor; here: b x h x w x D tokens)
                return {"patch_tokens": tok[:, 1:, :].reshape(B, hp, wp, cfg.dim)}
            mask_pred = torch.empty(B, L, cfg.n_queries, ho, wo, dtype=torch.float32, device=x.device)
            objectness = torch.empty(B, L, cfg.n_queries, dtype=torch.float32, device=x.device)
            features = torch.empty(B, cfg.dim, dtype=torch.float32, device=x.device)
            run(1 if L > 1 else 0, ptr(mask_pred), ptr(objectness), ptr(features))
        if self.return_intermediate:
            return {"objectness": objectness.unsqueeze(-1), "mask_pred": mask_pred, "features": features}
        return {"objectness": objectness[:, 0].unsqueeze(-1), "mask_pred": mask_pred[:, 0], "features": features}

    @torch.no_grad()
    def predict_one(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Single-image latency path (the consumer is `SelfMaskInference.predict`, app.py:241-347, which runs
        `base_structure._forward({'x': img})` on one image and takes `mask_pred[0, -1]` / `objectness[0, -1]`).

        x: 1 x 3 x H x W (float32 normalised, or uint8 raw pixels) on the GPU.  The ~170 launches of a batch-1 forward pass are
        latency-bound, so the pass is captured ONCE per (H, W, dtype) into a CUDA graph and replayed: the call costs one host →
        graph launch.  Returns the same dict as `forward` (tensors are the graph's static outputs: valid until the next call)."""
        if x.ndim != 4 or x.shape[0] != 1 or x.shape[1] != 3:
            raise SmkError(f"predict_one expects 1 x 3 x H x W, got {tuple(x.shape)}")
        _lib.require_cuda(x, "x")
        x = x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()
        key = (int(x.shape[2]), int(x.shape[3]), x.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            static_x = x.clone()
            with torch.cuda.device(x.device):
                side = torch.cuda.Stream(x.device)
                side.wait_stream(torch.cuda.current_stream(x.device))
                with torch.cuda.stream(side):          # warm-up outside capture: workspace, weight repack, layer-0 decoder constants
                    for _ in range(2):
                        self.forward(static_x)
                torch.cuda.current_stream(x.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self.forward(static_x)
            ent = (graph, static_x, static_out)
            self._graphs[key] = ent
        graph, static_x, static_out = ent
        static_x.copy_(x, non_blocking=True)
        graph.replay()
        return static_out

    def tap(self, what: int, B: int, H: int, W: int) -> torch.Tensor:
        """Stage-level parity hook: 1 final-LN tokens, 2 decoder queries [L,B,nq,D], 3 residual stream."""
        cfg = self.cfg
        handle = self._handles[(H, W)][0]
        N = (-(-H // cfg.patch)) * (-(-W // cfg.patch)) + 1
        shape = (cfg.dec_layers, B, cfg.n_queries, cfg.dim) if what == 2 else (B, N, cfg.dim)
        out = torch.empty(shape, dtype=torch.float32, device=self._blob.device)
        check(lib().smk_model_tap(handle, what, ptr(out), out.numel(), stream_ptr()), "smk_model_tap")
        return out


def get_model(arch: str, configs=None, **kwargs) -> SelfMaskB200:
    """Mirror of `utils/misc.py:163-188 get_model(arch="maskformer", configs=Namespace)`."""
    if arch != "maskformer":
        raise SmkError(f"{arch} is not on the B200 path; only arch='maskformer'")
    assert configs is not None
    return SelfMaskB200(n_queries=configs.n_queries, n_decoder_layers=configs.n_decoder_layers,
                        learnable_pixel_decoder=configs.learnable_pixel_decoder,
                        lateral_connection=configs.lateral_connection,
                        return_intermediate=configs.loss_every_decoder_layer, scale_factor=configs.scale_factor,
                        abs_2d_pe_init=configs.abs_2d_pe_init, use_binary_classifier=configs.use_binary_classifier,
                        arch=configs.arch, training_method=configs.training_method, patch_size=configs.patch_size, **kwargs)
