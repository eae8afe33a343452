#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of an object file / shared library (cuobjdump -sass): the table of
profiles/r0N_sass_excerpt.md.  usage: sass_table.py <file.o|.so> [name filter regex]"""
import re, subprocess, sys
path = sys.argv[1]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
dem = {}
rows = {}
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        r = rows[cur]
        r["n"] = r.get("n", 0) + 1
        for key, pat in (("UTMALDG", "UTMALDG"), ("UTMASTG", "UTMASTG"), ("SYNCS", "SYNCS"), ("F2", "FFMA2|FMUL2|FADD2"),
                         ("LDS", r"LDS"), ("STS", r"STS"), ("ATOMS", "ATOMS")):
            if re.match(pat, op):
                r[key] = r.get(key, 0) + 1
names = list(rows)
d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("| kernel | instructions | UTMALDG | UTMASTG | SYNCS | FFMA2+FMUL2+FADD2 | LDS | STS | ATOMS |")
print("|---|---|---|---|---|---|---|---|---|")
for mangled, name in sorted(zip(names, d), key=lambda t: t[1]):
    name = re.sub(r"^void mrla::", "", name)
    name = re.sub(r"\(CUtensorMap.*", "", name)
    name = name.replace("(bool)", "")
    if flt and not flt.search(name):
        continue
    r = rows[mangled]
    print(f"| `{name}` | {r.get('n', 0)} | " + " | ".join(str(r.get(k, 0)) for k in ("UTMALDG", "UTMASTG", "SYNCS", "F2", "LDS", "STS", "ATOMS")) + " |")
