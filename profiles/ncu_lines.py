#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
share of warp-stall samples and of executed warp instructions.  usage: ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
cur, hdr = None, None


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit() or len(r) < 10:
        continue
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    a = agg[(cur, int(r[0]))]
    a[0] += num(r[i_s])
    a[1] += num(r[i_i])
    a[2] += num(r[i_t])
    a[3] = r[1]
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("total samples", ts, "total warp inst", ti)
print("== by stall samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% thr/inst {a[2]/max(a[1],1):5.1f} | {a[3].strip()[:95]}")
print("== by instructions")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:4d} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% thr/inst {a[2]/max(a[1],1):5.1f} | {a[3].strip()[:95]}")
