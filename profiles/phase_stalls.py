#!/usr/bin/env python3
"""usage: phase_stalls.py src.csv  -- first kernel section of an `ncu --page source --csv --print-source sass`
export: stall samples per run of instructions with the same execution count (= per loop / phase), with the
dominant stall reasons and the hottest instructions of each run."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[heads[0]]
end = heads[1] - 1 if len(heads) > 1 else len(rows)
b = [r for r in rows[heads[0] + 1:end] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in b)
print(rows[0][1], 'instructions', len(b), 'samples', tot)
start = 0
for k in range(1, len(b) + 1):
    if k == len(b) or b[k][iex] != b[start][iex]:
        sm = sum(int(b[j][isamp]) for j in range(start, k))
        if sm >= 0.01 * tot:
            a = collections.Counter()
            for r in b[start:k]:
                for i in st:
                    a[hdr[i][6:]] += int(r[i] or 0)
            print(f"[{start:5d},{k:5d}) exec={b[start][iex]:>8s} samples={sm:6d} {100*sm/tot:5.1f}%  {[(n, v) for n, v in a.most_common(4)]}")
            hot = sorted(range(start, k), key=lambda j: -int(b[j][isamp]))[:4]
            for j in sorted(hot):
                if int(b[j][isamp]) >= 0.005 * tot:
                    print(f"        {j:5d} {100*int(b[j][isamp])/tot:5.1f}%  {b[j][isrc][:80]}")
        start = k
