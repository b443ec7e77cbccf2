#!/usr/bin/env python
"""Per-kernel SASS opcode evidence (B200_PROFILING.md "What proves a Blackwell-native kernel"): counts of the tensor-core / TMEM / TMA
mnemonics in every kernel of libselfmask_b200.so, from `cuobjdump -sass` — runs without a GPU.

    python scripts/sass_opcodes.py > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "salient-object-detection_b200", "libselfmask_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDGSTS", "MUFU", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    try:
        dem = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except Exception:
        demangle = {n: n for n in names}
    print("# SASS opcode counts per kernel of libselfmask_b200.so (cuobjdump -sass, sm_100a)")
    print("# UTC*MMA = tcgen05.mma | LDTM / STTM = tcgen05.ld / st (TMEM) | UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce")
    print("# HMMA = legacy mma.sync | LDGSTS = cp.async | MUFU = special-function unit | DFMA = fp64 FMA")
    hdr = f"{'kernel':80s} " + " ".join(f"{o:>8s}" for o in OPS) + f" {'instrs':>8s}"
    print(hdr)
    def short(n):
        d = demangle.get(n, n)
        if d.endswith(")"):                      # drop the trailing parameter list (template arguments keep their "(int)26" casts)
            depth = 0
            for i in range(len(d) - 1, -1, -1):
                depth += d[i] == ")"
                depth -= d[i] == "("
                if depth == 0:
                    d = d[:i]
                    break
        d = d.replace("(int)", "").replace("(bool)", "")
        d = d.replace("smk::", "").replace("(anonymous namespace)::", "")
        return d[:80]
    for n in sorted(names, key=lambda k: short(k)):
        c = counts[n]
        print(f"{short(n):80s} " + " ".join(f"{c[o]:8d}" for o in OPS) + f" {c['_total']:8d}")
    tc = [short(n) for n in names if counts[n]["UTCHMMA"] or counts[n]["UTCQMMA"]]
    hm = [short(n) for n in names if counts[n]["HMMA"] and not counts[n]["UTCHMMA"]]
    print(f"\n# kernels with tcgen05 MMAs: {len(tc)}; kernels with legacy HMMA only: {len(hm)}")
    for n in hm:
        print(f"#   HMMA-only: {n}")


if __name__ == "__main__":
    main()
