#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel share of one step
    python scripts/ncu_summary.py full gpurun_out/prof_gemm.ncu-rep [...]     # key metrics of every captured launch
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__cycles_active.avg", "cyc"),
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ix, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r[ix]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / tot:.3f} |")
    print(f"| **all** | {len(rows)} | {tot / 1e3:.1f} | 1.000 |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"### {path.split('/')[-1]}\n")
    print("| kernel | grid | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for d in data:
        cells = []
        for m, _ in KEYS:
            i = col.get(m)
            if i is None:
                cells.append("-")
                continue
            try:
                v = float(d[i].replace(",", ""))
                cells.append(f"{v:.4g} {units[i]}".strip())
            except ValueError:
                cells.append(d[i] or "-")
        print(f"| `{short(d[col['Kernel Name']])}` | {d[col['Grid Size']]} | " + " | ".join(cells) + " |")
    print()


def traffic(paths):
    """JSON for bench.py's roofline.traffic: DRAM bytes (read + write) per launch, keyed by stage, averaged over the
    captured launches weighted as they occur in one step (encoder-layer kernels x12, decoder-layer kernels x6)."""
    import json
    weight = {"enc": 12, "kv": 1, "dec": 6, "tail": 1}
    stage_of = [("gemm_bf16_tc", "gemm_tcgen05"), ("attn_tc", "attention_tcgen05"), ("layernorm", "layernorm"),
                ("query_iou", "eval_metrics"), ("mask_metrics", "eval_metrics"), ("mask_head", "mask_head"), ("mask_logits", "mask_head"),
                ("mask_upsample", "mask_head"), ("attn_small", "attention_tcgen05"), ("attn_fa", "attention_tcgen05")]
    acc = collections.OrderedDict()
    for path in paths:
        tag = next((k for k in weight if f"_{k}." in path), None)
        w = weight.get(tag, 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for d in data:
            name = short(d[col["Kernel Name"]])
            st = next((s_ for k, s_ in stage_of if k in name), "other")
            b = sum(float(d[col[m]].replace(",", "")) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            a = acc.setdefault(st, [0.0, 0])
            a[0] += w * b
            a[1] += w
    print(json.dumps({"source": [p_.split("/")[-1] for p_ in paths],
                      "note": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, launch-weighted mean per stage",
                      "dram_bytes_per_launch": {k: v[0] / v[1] for k, v in acc.items()},
                      "launches_weighted": {k: v[1] for k, v in acc.items()}}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
        sys.exit(0)
    kind, paths = sys.argv[1], sys.argv[2:]
    for p in paths:
        (launches if kind == "launches" else full)(p)
