#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per-instruction stall samples grouped into loop phases."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in rows[2:] if len(r) > isamp and r[isamp].isdigit())
print("total samples", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
acc = []
for n, r in enumerate(rows[2:]):
    if len(r) <= isamp or not r[isamp].isdigit():
        continue
    s = int(r[isamp])
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    acc.append((n, s, int(r[iex] or 0), r[isrc].strip(), st))
if top:
    for n, s, ex, src, st in sorted(acc, key=lambda x: -x[1])[:top]:
        print("%4d %6d %5.1f%% ex=%-9d %-60s %s" % (n, s, 100.0 * s / tot, ex, src[:60], st))
else:
    for n, s, ex, src, st in acc:
        print("%4d %6d %5.1f%% ex=%-9d %-70s %s" % (n, s, 100.0 * s / tot, ex, src[:70], st if s * 200 > tot else ""))
