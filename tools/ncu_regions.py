#!/usr/bin/env python
"""Instruction and stall-sample shares per (file, line range) of an ncu source page:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv; python tools/ncu_regions.py f.csv file:lo-hi=name ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for spec in sys.argv[2:]:
    loc, name = spec.split("=")
    f, rng = loc.split(":")
    lo, hi = map(int, rng.split("-"))
    regions.append((f, lo, hi, name))
cur = "?"; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples"); continue
    if hdr and len(r) > ii:
        try: data.append((cur, int(r[0]), int(r[ii]), int(r[sa])))
        except ValueError: pass
tot = sum(d[2] for d in data); ts = sum(d[3] for d in data)
acc = {}
for f, ln, n, smp in data:
    key = f
    for rf, lo, hi, name in regions:
        if rf == f and lo <= ln <= hi: key = name; break
    a = acc.setdefault(key, [0, 0]); a[0] += n; a[1] += smp
print(f"total warp-instructions {tot}, samples {ts}")
for k, (n, smp) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*n/tot:6.1f} % inst  {100*smp/ts:6.1f} % samples  {k}")
