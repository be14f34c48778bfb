#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an ncu report: `ncu -i X.ncu-rep --page source --csv | ncu_hot.py`."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = None
out = []
kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr and len(r) > 10:
        out.append(r)
tot = sum(int(r[hdr["# Samples"]]) for r in out) or 1
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[hdr[s]]) for r in out) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(range(len(out)), key=lambda i: -int(out[i][hdr["# Samples"]]))[: int(sys.argv[1]) if len(sys.argv) > 1 else 25]
for i in sorted(top):
    r = out[i]
    main = max(stalls, key=lambda s: int(r[hdr[s]]))
    print(f"{i:5d} {100*int(r[hdr['# Samples']])/tot:5.1f}%  {main:<18} {r[hdr['Source']].strip()[:90]}")
