"""CPU emulation of tensor-core operand precision schedules for the encoder (design aid, not product code).

Every encoder contraction (patch-embed, qkv, Q.K^T, P.V, proj, fc1, fc2) and the decoder-memory K/V projection is
computed with its operands rounded as a given numeric scheme would round them, fp32 accumulate; LayerNorm / softmax /
residual / GELU stay fp32; decoder, objectness and mask contraction stay fp32 (the shipped modes run them as 3-term
splits).  Output: the north_star parity criteria (logits max-abs over all decoder layers at mask resolution, mean / min
binarised-IoU agreement of the last layer, objectness top-1) against the pure-fp32 oracle.

schemes per GEMM:  f32 | b1 (bf16 x bf16) | h1 (fp16 x fp16) | h2a (A = hi+lo fp16, W fp16) | h2w (A fp16, W = hi+lo)
                   | h3 / b3 (3-term split)

usage: python scripts/precision_emulation.py "h1" "h3:0-3,h1" ...   (schedule: default scheme, or scheme:block-range lists)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import selfmask_oracle as O  # noqa: E402


def rnd(x, dt):
    return x.to(dt).to(torch.float32)


def split(x, dt):
    hi = rnd(x, dt)
    return hi, rnd(x - hi, dt)


def mm(a, w_t, scheme):
    """a [.., K] @ w_t [K, N] with operand rounding per scheme."""
    if scheme == "f32":
        return a @ w_t
    dt = torch.bfloat16 if scheme[0] == "b" else torch.float16
    kind = scheme[1:]
    if kind == "1":
        return rnd(a, dt) @ rnd(w_t, dt)
    ah, al = split(a, dt)
    wh, wl = split(w_t, dt)
    if kind == "3q":
        # fp16 main term + the two correction terms on e4m3 operands (fp8 tensor-core rate): A_hi8·(W_lo·2^15)8 + (A_lo·2^11)8·(W_hi·2^4)8,
        # accumulated first and scaled by 2^-15 when the fp16 term joins (tcgen05 scale-input-d)
        q8 = lambda t: t.clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32)
        corr = q8(ah) @ q8(wl * 2.0 ** 15) + q8(al * 2.0 ** 11) @ q8(wh * 2.0 ** 4)
        return ah @ wh + corr * 2.0 ** -15
    if kind == "2a":
        return ah @ wh + al @ wh
    if kind == "2w":
        return ah @ wh + ah @ wl
    if kind == "3":
        return ah @ wh + (ah @ wl + al @ wh)
    raise ValueError(scheme)


def encoder(sd, x, cfg, sched):
    P, D, H = cfg["patch_size"], cfg["dim"], cfg["heads"]
    dh = D // H
    b, _, h0, w0 = x.shape
    hp, wp = h0 // P, w0 // P
    cols = x.reshape(b, 3, hp, P, wp, P).permute(0, 2, 4, 1, 3, 5).reshape(b, hp * wp, 3 * P * P)
    wmat = sd["encoder.patch_embed.proj.weight"].reshape(D, 3 * P * P)
    tok = mm(cols, wmat.t(), sched["pe"]) + sd["encoder.patch_embed.proj.bias"]
    t = torch.cat((sd["encoder.cls_token"].expand(b, -1, -1), tok), dim=1) + O.interpolate_pos_embed(sd["encoder.pos_embed"], hp, wp)
    for i in range(cfg["depth"]):
        s = sched["blocks"][i]
        p = f"encoder.blocks.{i}."
        n = t.shape[1]
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = (mm(y, sd[p + "attn.qkv.weight"].t(), s["qkv"]) + sd[p + "attn.qkv.bias"]).reshape(b, n, 3, H, dh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (mm(q, k.transpose(-2, -1), s["qk"]) * dh ** -0.5).softmax(dim=-1)
        y = mm(attn, v, s["pv"]).transpose(1, 2).reshape(b, n, D)
        t = t + (mm(y, sd[p + "attn.proj.weight"].t(), s["proj"]) + sd[p + "attn.proj.bias"])
        y = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.gelu(mm(y, sd[p + "mlp.fc1.weight"].t(), s["fc1"]) + sd[p + "mlp.fc1.bias"])
        t = t + (mm(y, sd[p + "mlp.fc2.weight"].t(), s["fc2"]) + sd[p + "mlp.fc2.bias"])
    return F.layer_norm(t, (D,), sd["encoder.norm.weight"], sd["encoder.norm.bias"], 1e-6)


def decoder(sd, memory, cfg, kv_scheme, da="f32"):
    """oracle decoder with the memory K/V projection under `kv_scheme`, the rest fp32."""
    D, H = cfg["dim"], cfg["heads"]
    dh = D // H
    b = memory.shape[0]
    qpos = sd["query_embed"].unsqueeze(0).expand(b, -1, -1)
    tgt = torch.zeros_like(qpos)
    inter = []
    for i in range(cfg["n_decoder_layers"]):
        p = f"decoder.layers.{i}."
        qk = tgt + qpos
        tgt = tgt + _mha_emu(sd, p + "self_attn", qk, qk, tgt, H, da)
        tgt = F.layer_norm(tgt, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        w, bias = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]
        q = (tgt + qpos) @ w[:D].t() + bias[:D]
        k = mm(memory, w[D:2 * D].t(), kv_scheme) + bias[D:2 * D]
        v = mm(memory, w[2 * D:].t(), kv_scheme) + bias[2 * D:]
        lq, lk = q.shape[1], k.shape[1]
        q = q.reshape(b, lq, H, dh).transpose(1, 2) * dh ** -0.5
        k = k.reshape(b, lk, H, dh).transpose(1, 2)
        v = v.reshape(b, lk, H, dh).transpose(1, 2)
        if da != "f32":      # q, k, v are stored in the attention operand type (q after the dh^-0.5 scale is folded into the kernel)
            dt = torch.bfloat16 if da[0] == "b" else torch.float16
            k, v = rnd(k, dt), rnd(v, dt)
            a = (rnd(q / dh ** -0.5, dt) * dh ** -0.5 @ k.transpose(-2, -1)).softmax(dim=-1)
            o = (rnd(a, dt) @ v).transpose(1, 2).reshape(b, lq, D)
        else:
            a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
            o = (a @ v).transpose(1, 2).reshape(b, lq, D)
        tgt = tgt + (o @ sd[p + "multihead_attn.out_proj.weight"].t() + sd[p + "multihead_attn.out_proj.bias"])
        tgt = F.layer_norm(tgt, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        y = F.relu(tgt @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
        tgt = tgt + (y @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"])
        tgt = F.layer_norm(tgt, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
        inter.append(F.layer_norm(tgt, (D,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], 1e-5))
    return torch.stack(inter, dim=1)


def _mha_emu(sd, prefix, query, key, value, heads, da):
    if da == "f32":
        return O._mha(sd, prefix, query, key, value, heads)
    dt = torch.bfloat16 if da[0] == "b" else torch.float16
    D = query.shape[-1]
    dh = D // heads
    w, bias = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = rnd(query @ w[:D].t() + bias[:D], dt)
    k = rnd(key @ w[D:2 * D].t() + bias[D:2 * D], dt)
    v = rnd(value @ w[2 * D:].t() + bias[2 * D:], dt)
    b, lq, _ = q.shape
    lk = k.shape[1]
    q = q.reshape(b, lq, heads, dh).transpose(1, 2) * dh ** -0.5
    k = k.reshape(b, lk, heads, dh).transpose(1, 2)
    v = v.reshape(b, lk, heads, dh).transpose(1, 2)
    a = rnd((q @ k.transpose(-2, -1)).softmax(dim=-1), dt)
    o = (a @ v).transpose(1, 2).reshape(b, lq, D)
    return o @ sd[prefix + ".out_proj.weight"].t() + sd[prefix + ".out_proj.bias"]


GEMMS = ("qkv", "qk", "pv", "proj", "fc1", "fc2")


def parse(spec, depth):
    """'h1' | 'h3:0-3,h1' | 'h3:0-1,h2a:2-5,h1' ; optional per-GEMM override 'h1/qk=h3/pv=h3'; 'pe=..' and 'kv=..' likewise."""
    parts = spec.split("/")
    sched = {"pe": None, "kv": None, "da": "f32", "blocks": [dict() for _ in range(depth)]}
    default = None
    for item in parts[0].split(","):
        if ":" in item:
            sc, rng = item.split(":")
            a, _, b_ = rng.partition("-")
            for i in range(int(a), int(b_ or a) + 1):
                for g in GEMMS:
                    sched["blocks"][i][g] = sc
        else:
            default = item
    for blk in sched["blocks"]:
        for g in GEMMS:
            blk.setdefault(g, default)
    sched["pe"] = sched["blocks"][0]["qkv"]
    sched["kv"] = default
    for ov in parts[1:]:
        k, v = ov.split("=")
        if k in ("pe", "kv", "da"):
            sched[k] = v
        else:
            for blk in sched["blocks"]:
                blk[k] = v
    return sched


def main():
    specs = sys.argv[1:] or ["f32", "b1", "h1"]
    nq, B, H, W = 20, int(os.environ.get("EMU_B", 8)), int(os.environ.get("EMU_HW", 224)), int(os.environ.get("EMU_HW", 224))
    cfg = O.make_config(n_queries=nq)
    sd = O.synth_state_dict(cfg, seed=int(os.environ.get("EMU_WSEED", 0)))
    x = O.normalize_images(O.synth_images_u8(B, H, W, seed=int(os.environ.get("EMU_SEED", 99))))
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = O.model_forward(sd, x, cfg, return_logits=True)
        for spec in specs:
            sched = parse(spec, cfg["depth"])
            tokens = encoder(sd, x, cfg, sched)
            memory = tokens[:, 1:, :]
            queries = decoder(sd, memory, cfg, sched["kv"], sched["da"])
            hp = H // cfg["patch_size"]
            feat = memory.transpose(1, 2).reshape(B, cfg["dim"], hp, hp)
            up = F.interpolate(feat, scale_factor=cfg["scale_factor"], mode="bilinear")
            logits = torch.einsum("bdqn,bnhw->bdqhw", queries, up)
            obj = O.objectness_head(sd, queries)
            pa = torch.sigmoid(logits[:, -1]).numpy() > 0.5
            pb = ref["mask_pred"][:, -1].numpy() > 0.5
            inter = (pa & pb).sum((-1, -2)).astype(np.float64)
            union = (pa | pb).sum((-1, -2)).astype(np.float64)
            agree = np.where(union == 0, 1.0, inter / np.maximum(union, 1))
            top = int((obj[:, -1, :, 0].argmax(-1) == ref["objectness"][:, -1, :, 0].argmax(-1)).sum())
            d = (logits - ref["mask_logits"]).abs()
            print(f"{spec:40s} logits max-abs {float(d.max()):.4f} (last {float(d[:, -1].max()):.4f})  tok {float((tokens - ref['tokens']).abs().max()):.2e}"
                  f"  IoU agree mean {agree.mean() * 100:.3f}% min {agree.min() * 100:.2f}%  top1 {top}/{B}", flush=True)


if __name__ == "__main__":
    main()
