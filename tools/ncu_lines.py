#!/usr/bin/env python
"""Per-source-line summary of an ncu report: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv;
python tools/ncu_lines.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = "?"
data = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
        sb = hdr.index("stall_barrier"); ss = hdr.index("stall_short_sb"); sw = hdr.index("stall_wait")
        sl = hdr.index("stall_long_sb"); sm = hdr.index("stall_math"); sn = hdr.index("stall_no_inst")
        continue
    if hdr and len(r) > ii and r[0] not in ("", "Line No"):
        try:
            data.append((int(r[ii]), int(r[sa]), cur_file, r[0], r[1].strip()[:90],
                         int(r[sb] or 0), int(r[ss] or 0), int(r[sw] or 0), int(r[sl] or 0), int(r[sm] or 0), int(r[sn] or 0)))
        except ValueError:
            pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"total warp-instructions {tot}  samples {ts}")
print(" inst%  smp%  [barrier short_sb wait long_sb math no_inst]  file:line  source")
key = (lambda d: d[1]) if len(sys.argv) > 3 else (lambda d: d[0])
for d in sorted(data, key=key, reverse=True)[:top]:
    print(f"{d[0]/tot*100:5.1f} {d[1]/ts*100:5.1f}  [{d[5]:5d} {d[6]:5d} {d[7]:5d} {d[8]:5d} {d[9]:5d} {d[10]:5d}]  {d[2]}:{d[3]}  {d[4]}")
