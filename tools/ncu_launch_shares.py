#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share."""
import collections
import csv
import re
import sys


def main(path, top=45):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    n_launch = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", re.sub(r"void mrla::|mrla::|void ", "", row["Kernel Name"]))[:96]
        unit = row.get("Metric Unit", "ns")
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        n_launch += 1
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("k_"))
    grp = lambda pre: sum(a[1] for k, a in agg.items() if k.startswith(pre))
    print(f"* launches: {n_launch}; device time: {tot / 1e3:.2f} ms")
    print(f"* this repo's kernels (`k_*`, namespace mrla): {ours / 1e3:.2f} ms = {100 * ours / tot:.1f} % — MRLA tail group "
          f"(`k_light_*`) {grp('k_light') / 1e3:.2f} ms = {100 * grp('k_light') / tot:.1f} %, BatchNorm op (`k_bn_*`) "
          f"{grp('k_bn') / 1e3:.2f} ms = {100 * grp('k_bn') / tot:.1f} %, stem max pooling (`k_maxpool*`) "
          f"{grp('k_maxpool') / 1e3:.2f} ms = {100 * grp('k_maxpool') / tot:.1f} %")
    print()
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {k} | {n} | {us:.0f} | {100 * us / tot:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
