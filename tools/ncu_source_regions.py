#!/usr/bin/env python
"""Group an `ncu --page source --csv` dump by per-instruction execution count (= loop region) and show
instruction mix and stall samples per region, then the top stalled instructions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
want = sys.argv[3] if len(sys.argv) > 3 else None   # optional kernel-name substring
sel = 0
if want:
    for n, i in enumerate(hdr_idx):
        if want in ' '.join(rows[i - 1]):
            sel = n
            break
hi = hdr_idx[sel]
end = hdr_idx[sel + 1] - 1 if len(hdr_idx) > sel + 1 else len(rows)
print('kernel:', ' '.join(rows[hi - 1])[:120])
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si = hdr.index('Warp Stall Sampling (All Samples)'); ii = hdr.index('Instructions Executed'); so = hdr.index('Source')
tot = sum(int(r[si]) for r in data); toti = sum(int(r[ii]) for r in data)
print(len(data), 'SASS instructions; total samples', tot, 'total warp-inst', toti)


def op(s):
    t = s.split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]


b = collections.Counter(); bs = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in data:
    c = int(r[ii]); b[c] += c; bs[c] += int(r[si]); ops[c][op(r[so])] += 1
for k in sorted(b, key=lambda k: -b[k])[:10]:
    print(f"exec/inst={k:10d} inst={b[k]:12d} ({b[k]/toti*100:5.1f}%) samples={bs[k]:7d} ({bs[k]/tot*100:5.1f}%)", dict(ops[k].most_common(10)))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
print('top stalled instructions:')
for r in sorted(data, key=lambda r: -int(r[si]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    reasons = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i] not in ('', '0')), reverse=True)[:3]
    print(f"{int(r[si]):6d} {int(r[ii]):10d}  {r[so][:70]:70s} {reasons}")
