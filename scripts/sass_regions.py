#!/usr/bin/env python
"""Per-region executed-instruction counts from an `ncu --page source --csv --print-source sass` dump.
usage: sass_regions.py file.csv events_per_launch [min_instr_per_event]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ev = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 15
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ia]) for r in data)
print("instructions", len(data), "executed/event", round(tot / ev, 1), "samples", sum(int(r[ismp]) for r in data))
cur = None; start = 0; acc = 0; smp = 0; out = []
for i, r in enumerate(data):
    c = int(r[ia])
    if cur is None: cur = c; start = i
    if abs(c - cur) > 0.02 * max(cur, 1):
        out.append((start, i - 1, cur, acc, smp)); cur = c; start = i; acc = 0; smp = 0
    acc += c; smp += int(r[ismp])
out.append((start, len(data) - 1, cur, acc, smp))
for s, e, c, a, sm in out:
    if a / ev > thr: print(f"{s:5d}-{e:5d} n={e-s+1:4d} exec/ev={c/ev:6.2f} instr/ev={a/ev:7.1f} samples={sm}  | {data[s][isrc].strip()[:60]}")
