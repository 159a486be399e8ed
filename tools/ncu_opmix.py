#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode (executed warp-instructions)."""
import csv, collections, sys
path, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]
ci, ei, si = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot, samp = collections.Counter(), collections.Counter()
cnt = 0
for r in rows[2:]:
    try: n = int(r[ei])
    except Exception: continue
    op = r[ci].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith('@') else op[0]
    o = '.'.join(o.split('.')[:2]) if o.startswith(('IMAD', 'MUFU', 'LDG', 'STG', 'I2F', 'F2F', 'SHF')) else o.split('.')[0]
    tot[o] += n; cnt += n; samp[o] += int(r[si] or 0)
print(f"total warp-inst {cnt}  = {cnt/units:.1f} per unit")
for o, n in tot.most_common(45):
    print(f"{o:16s} {n:10d} {n/units:7.2f}/unit  samples {samp[o]}")
