#!/usr/bin/env python
"""Executed warp-instructions per address range of an `ncu --page source --csv` dump.
usage: ncu_ranges.py src.csv lo:hi[:divisor] ...   (hex addresses as printed in the CSV's last 5 hex digits)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, ei, ai, si = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Address'), hdr.index('# Samples')
data = []
for r in rows[2:]:
    try: data.append((int(r[ai], 16) & 0xFFFFF, int(r[ei]), r[ci].strip(), int(r[si] or 0)))
    except Exception: pass
base = data[0][0]
for spec in sys.argv[2:]:
    p = spec.split(':')
    lo, hi = int(p[0], 16), int(p[1], 16)
    div = float(p[2]) if len(p) > 2 else 1.0
    tot = collections.Counter(); n = 0; smp = 0
    for a, e, s, sm in data:
        if lo <= a - base <= hi:
            op = s.split()
            o = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
            o = '.'.join(o.split('.')[:2]) if o.startswith(('IMAD', 'MUFU', 'LDG', 'STG', 'LDS', 'STS')) else o.split('.')[0]
            tot[o] += e; n += e; smp += sm
    print(f"range {spec}: executed {n} = {n/div:.1f} per unit, samples {smp}")
    print('   ' + ', '.join(f"{o} {c/div:.1f}" for o, c in tot.most_common(28)))
