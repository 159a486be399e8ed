import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import stochquant_b200 as sq
dims = (64, 64, 64, 64)
V = int(np.prod(dims))
for flags in (0, sq.SQ_FLAG_NO_OBSERVABLES):
    ctx = sq.Context(dims, real="f32", math="fast", flags=flags)
    for _ in range(3):
        ctx.step(0.01, 100)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); ctx.step(0.01, 100); ts.append(time.perf_counter() - t0)
    dt = np.median(ts)
    print(f"SQ_PDL={os.environ.get('SQ_PDL')} flags={flags}: {dt*1e4:.2f} us per tau-step, {V*100/dt/1e9:.1f} G/s")
    ctx.close()
