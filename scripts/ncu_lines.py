#!/usr/bin/env python
"""Rank CUDA source lines of an ncu report by executed instructions / stall samples.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; ncu_lines.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; out = []
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if not r or r[0] in ('Line No', '', 'Function Name'): continue
    if len(r) > 8 and r[0].isdigit() and r[2] == '-':
        out.append((cur, int(r[0]), r[1].strip()[:100], num(r[4]), num(r[7])))
ts = sum(o[3] for o in out) or 1; ti = sum(o[4] for o in out) or 1
print("total samples", ts, "total warp-inst", ti)
agg = {}
for o in out:
    k = (o[0], o[1]); a = agg.setdefault(k, [o[2], 0, 0]); a[1] += o[3]; a[2] += o[4]
items = sorted(agg.items(), key=lambda kv: -kv[1][2])
for (f, ln), (src, s, i) in items[:N]:
    print(f"{f:22s} {ln:4d} inst {100*i/ti:5.1f}% samp {100*s/ts:5.1f}%  {src}")
