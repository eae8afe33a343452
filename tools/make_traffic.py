#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full --page raw --csv` export of ONE forward+backward of the stage-1 tail op
(tools/final_profile_r02.sh): sum of dram__bytes_read.sum + dram__bytes_write.sum over the op's kernels, per kernel too."""
import csv, json, subprocess, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def val(r, m):
    return float(r[col[m]].replace(",", "")) * scale.get(units[col[m]], 1)
per, tot = [], 0.0
for r in data:
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    us = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    u = units[col["gpu__time_duration.sum"]]
    us = us * (1e3 if u in ("ms", "msecond") else (1e-3 if u in ("ns", "nsecond") else 1))
    per.append({"kernel": r[col["Kernel Name"]][:60], "dram_bytes": int(b), "us": round(us, 1)})
    tot += b
try:
    commit = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    commit = sys.argv[3] if len(sys.argv) > 3 else "unknown"
out = {"stage1_tail_fwd_bwd_dram_bytes": int(tot), "measured_at_commit": commit, "file": "profiles/r02_final_tail_stage1_full.md",
       "how": "ncu --set full --clock-control none, one launch of each kernel of the op (tools/final_profile_r02.sh), cold L2",
       "algorithmic_bytes_8N": 8 * 256 * 256 * 56 * 56 * 2, "kernels": per}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
