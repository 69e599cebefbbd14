#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: python tools/ncu_lines.py file.csv [top]"""
import csv, sys
def num(x):
    try: return int(x)
    except Exception: return 0
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
ci = {n: i for i, n in reversed(list(enumerate(H)))}
iw, it, isamp = ci["Instructions Executed"], ci["Thread Instructions Executed"], ci["# Samples"]
lines = []
for r in rows[hdr + 1:]:
    if r and r[0].strip().isdigit():
        lines.append((int(r[0]), r[1].strip(), num(r[iw]), num(r[it]), num(r[isamp])))
tw = sum(l[2] for l in lines); ts = sum(l[4] for l in lines)
print("total warp inst %d, thread inst %d, samples %d" % (tw, sum(l[3] for l in lines), ts))
for ln, src, w, t, s in sorted(lines, key=lambda l: -l[4])[:top]:
    print("%5d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f  %s" % (ln, 100.0 * w / tw, 100.0 * s / max(ts, 1), t / max(w, 1), src[:150]))
