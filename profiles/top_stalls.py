#!/usr/bin/env python3
"""usage: top_stalls.py src.csv [N]  -- the N SASS instructions with the most stall samples of an
`ncu --page source --csv --print-source sass` export, with the dominant stall reasons of each."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
agg = {}
for r in body:
    for i in st:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: round(100 * v / max(1, sum(agg.values())), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:N]
for k in sorted(order):
    r = body[k]
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in st), reverse=True)[:3]
    print(f"{k:5d} {100*int(r[isamp])/tot:5.2f}%  {r[isrc][:70]:70s} {[(n, v) for v, n in reasons if v]}")
