#!/usr/bin/env python
"""List backward branches (loops) of a cuobjdump -sass dump and the opcode mix of a chosen range.
usage: sass_loops.py file.sass [lo hi]   (hex addresses)"""
import re, sys, collections
ins = []
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print('total instructions', len(ins), 'last addr', hex(ins[-1][0]))
if len(sys.argv) < 4:
    for a, t in ins:
        if 'BRA' in t:
            mm = re.search(r'0x([0-9a-f]+)', t)
            if mm and int(mm.group(1), 16) < a: print(hex(a), '->', hex(int(mm.group(1), 16)), 'span', (a - int(mm.group(1), 16)) // 16, 'instr |', t[:70])
else:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    c = collections.Counter()
    n = 0
    for a, t in ins:
        if lo <= a <= hi:
            op = t.split()
            o = op[1] if op[0].startswith('@') else op[0]
            o = '.'.join(o.split('.')[:2]) if o.startswith(('IMAD', 'MUFU', 'LDG', 'STG', 'I2F', 'F2F', 'LDS', 'STS', 'LDL', 'STL')) else o.split('.')[0]
            c[o] += 1; n += 1
    print('range instr', n)
    for o, k in c.most_common(50): print(f'{o:14s} {k}')
