#!/usr/bin/env python
"""Top stall-sample SASS lines + headline metrics of an .ncu-rep (first kernel)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, d = rows[0], rows[1], rows[2]
for w in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']:
    if w in hdr: print(f"{w:66s} {units[hdr.index(w)]:10s} {d[hdr.index(w)]}")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and 'per_issue_active' in h and float(d[i] or 0) > 0.2: print(f"  {h.split('issue_stalled_')[1].split('_per')[0]:22s} {float(d[i]):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ci, si, ei, ai = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Address')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []; tot = 0
for r in rows[2:]:
    try: n = int(r[si])
    except Exception: continue
    tot += n
    top = max(stall_cols, key=lambda i: int(r[i] or 0))
    data.append((n, r[ai][-6:], r[ci].strip(), r[ei], hdr[top]))
print('total samples', tot)
for n, a, s, e, st in sorted(data, reverse=True)[:topn]:
    print(f"{n:7d} {100*n/tot:5.1f}% {a} exec={e:>8s} {st:18s} {s[:80]}")
