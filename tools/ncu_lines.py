#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line of the first kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  usage: ncu_lines.py rep units [topn]"""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, data = None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] == 'Line No': hdr = r; ei = r.index('Instructions Executed'); continue
    if hdr and len(r) > 8 and r[0] != '':
        try: n = int(r[ei])
        except Exception: continue
        if n:
            k = (cur, int(r[0]), r[1].strip()[:100]); data[k] = data.get(k, 0) + n
print('total per unit', sum(data.values()) / units)
for (f, l, sx), n in sorted(data.items(), key=lambda x: -x[1])[:topn]: print(f"{n/units:6.2f} {f}:{l}: {sx}")
