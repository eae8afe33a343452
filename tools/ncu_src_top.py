#!/usr/bin/env python
"""Top stalled SASS instructions of an `ncu --page source --csv` export (sampling counts, dominant stall reason).
The export may hold several kernels, each with its own header row; every section is summarised."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur = [], None
for r in rows:
    if "# Samples" in r and "Source" in r:
        cur = {"hdr": r, "rows": [], "title": sections[-1]["pending"] if sections and "pending" in sections[-1] else ""}
        sections.append(cur)
    elif cur is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
    elif len(r) >= 1 and cur is None:
        pass
for sec in sections:
    hdr = sec["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for idx, r in enumerate(sec["rows"]):
        try:
            n = int(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        ex = r[col["Instructions Executed"]] if "Instructions Executed" in col else "0"
        data.append((n, idx, r[col["Source"]].strip(), int(ex or 0), st))
    tot = sum(d[0] for d in data) or 1
    print("== section: total samples", tot, "instructions", len(data))
    agg = {}
    for h in stall_cols:
        s = 0
        for r in sec["rows"]:
            try:
                s += int(r[col[h]] or 0)
            except ValueError:
                pass
        agg[h[6:]] = s
    print("by reason:", ", ".join(f"{k} {v * 100 // tot}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for n, idx, src, ex, st in sorted(data, reverse=True)[:top]:
        print(f"{n:6d} {100 * n / tot:5.1f}%  #{idx:4d} x{ex:8d}  {src[:86]:86s} " + " ".join(f"{k}:{v}" for v, k in st if v))
