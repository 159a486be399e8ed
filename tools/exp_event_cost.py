"""Experiment: what does one RNG event cost a 1000-step frame of 1024^2 (on-chip kernel)?  Wall time per frame (host clock
around sq_step, which synchronises) against the number of events the frame replayed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stochquant_b200 as sq
ctx = sq.Context((1024, 1024), real="f32", math="fast", seed=1242608872)
for _ in range(5):
    ctx.step(0.01, 1000)
rows, ne = [], ctx.measure()["nevents"]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    t0 = time.perf_counter()
    ctx.step(0.01, 1000)
    dt = (time.perf_counter() - t0) * 1e3
    n2 = ctx.measure()["nevents"]
    rows.append((dt, n2 - ne))
    ne = n2
a = np.array(rows)
for k in sorted(set(a[:, 1].astype(int))):
    s = a[a[:, 1] == k, 0]
    print(f"frames with {k} event(s): n={len(s):3d}  median {np.median(s):.3f} ms  min {s.min():.3f}  max {s.max():.3f}")
base = np.median(a[a[:, 1] == 0, 0])
ev = a[a[:, 1] > 0]
if len(ev):
    print(f"cost per event: {np.mean((ev[:, 0] - base) / ev[:, 1]):.3f} ms (mean over {len(ev)} frames); events per frame {a[:, 1].mean():.2f}")
