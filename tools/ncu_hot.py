#!/usr/bin/env python
"""ncu_hot.py -- per-SASS-instruction counts and stall samples of a kernel from `ncu --page source --csv` output.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [min_executed]
Prints address offset, instructions executed (warp level), stall samples, dominant stall reason, SASS text.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
tot_ex = sum(int(r[iex]) for r in rows[2:] if r[iex].isdigit())
tot_s = sum(int(r[ismp]) for r in rows[2:] if r[ismp].isdigit())
print(f"total executed {tot_ex}, samples {tot_s}")
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[2:]:
    ex, smp = int(r[iex]), int(r[ismp])
    if ex < mn:
        continue
    st = max(stall_cols, key=lambda c: int(r[c[0]] or 0))
    print(f"{int(r[ia], 16) - base:6x} {ex:10d} {100.0 * ex / tot_ex:5.2f}% smp {smp:6d} {100.0 * smp / max(tot_s, 1):5.2f}% {st[1][6:]:>12s} {r[isrc].strip()}")
