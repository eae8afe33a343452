#!/usr/bin/env python
"""Top stalled SASS instructions of an `ncu --page source --csv` export (sampling counts, dominant stall reason)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    data.append((n, idx, r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0), st))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for h in stall_cols:
    agg[h[6:]] = sum(int(r[col[h]] or 0) for r in rows[2:] if len(r) >= len(hdr))
print("by reason:", ", ".join(f"{k} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, idx, src, ex, st in sorted(data, reverse=True)[:top]:
    print(f"{n:6d} {100 * n / tot:5.1f}%  #{idx:4d} x{ex:8d}  {src[:70]:70s} " + " ".join(f"{k}:{v}" for v, k in st if v))
