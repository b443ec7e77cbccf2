#!/usr/bin/env python
"""Top stall PCs of one kernel from an ncu report with source (tuning aid).
Usage: python scripts/ncu_hot.py report.ncu-rep kernel_regex [n_top]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n_top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not hi:
    sys.exit("no kernel matched")
hdr = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
data = [r for r in rows[hi[0] + 1:end] if len(r) >= len(hdr)]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot, n, ins = collections.Counter(), 0, 0
for r in data:
    n += int(r[iN] or 0)
    ins += int(r[iI] or 0)
    for i in stall:
        tot[hdr[i][6:]] += int(r[i] or 0)
print(rows[hi[0] - 1][1][:100] if hi[0] > 0 else "")
print("samples", n, "warp instructions", ins)
print(tot.most_common(10))
for r in sorted(data, key=lambda r: -int(r[iN] or 0))[:n_top]:
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
    print(r[iN].rjust(6), r[iI].rjust(9), r[iS][:72].ljust(72), dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
ops = collections.Counter()
for r in data:
    t = r[iS].split()
    if t:
        ops[(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += int(r[iI] or 0)
print(ops.most_common(20))
