#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (per-launch metrics) into one line per distinct kernel/grid."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    data = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], row["Kernel Name"], row["Grid Size"], row["Block Size"])
        data.setdefault(key, {})[row["Metric Name"]] = row["Metric Value"]
    agg = collections.OrderedDict()
    for (_id, name, grid, block), v in data.items():
        name = re.sub(r"void mrla::|mrla::", "", name)
        name = re.sub(r"\(.*", "", name)
        sig = (name, grid, block)
        a = agg.setdefault(sig, collections.defaultdict(list))
        for m, val in v.items():
            try:
                a[m].append(float(val.replace(",", "")))
            except ValueError:
                pass
    rows = []
    for (name, grid, block), a in agg.items():
        n = len(a["gpu__time_duration.sum"])
        us = sum(a["gpu__time_duration.sum"]) / n / 1e3
        r = [name[:64], grid, block, n, f"{us:.1f}"]
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            r.append(f"{sum(a[m]) / len(a[m]) / 1e6:.0f}" if a.get(m) else "-")
        for m in ("smsp__inst_executed.sum",):
            r.append(f"{sum(a[m]) / len(a[m]) / 1e6:.1f}" if a.get(m) else "-")
        for m in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
            r.append(f"{sum(a[m]) / len(a[m]):.1f}" if a.get(m) else "-")
        rows.append(r)
    hdr = ["kernel", "grid", "block", "launches", "avg_us", "dram_rd_MB", "dram_wr_MB", "warp_inst_M", "issue_pct", "dram_pct"]
    text = "| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr) + "\n" + "\n".join("| " + " | ".join(map(str, r)) + " |" for r in rows)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
