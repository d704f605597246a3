#!/usr/bin/env python
"""Aggregate the ncu source page (SASS) of a report into chunks: samples, instructions, dominant opcodes, top stalls.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_regions.py src.csv [chunk]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = rows[1]
iS, iI, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx = {h: hdr.index(h) for h in names}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 5]
tot = sum(int(r[iS]) for r in data); toti = sum(int(r[iI]) for r in data)
print('total samples', tot, 'total warp-instr', toti, 'SASS lines', len(data))
def op(r):
    t = r[iSrc].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for s in range(0, len(data), chunk):
    seg = data[s:s + chunk]
    smp = sum(int(r[iS]) for r in seg); ins = sum(int(r[iI]) for r in seg)
    if smp < tot * 0.01 and ins < toti * 0.01:
        continue
    ops = collections.Counter(op(r) for r in seg).most_common(5)
    st = collections.Counter()
    for r in seg:
        for h in names:
            st[h] += int(r[idx[h]] or 0)
    print(f"[{s:4d}-{s+chunk:4d}] samples {100*smp/tot:5.1f}%  instr {100*ins/toti:5.1f}%  ops {ops}  stalls {[(k[6:], v) for k, v in st.most_common(4)]}")
