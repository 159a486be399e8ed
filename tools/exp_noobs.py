"""Experiment: per-step cost of the separate finalize launch on 64^4 (observables on/off)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochquant_b200 as sq, numpy as np
dims = (64, 64, 64, 64); V = int(np.prod(dims))
for flags in (0, 1):
    ctx = sq.Context(dims, real="f32", math="fast", flags=flags)
    for _ in range(3): ctx.step(0.01, 100)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); ctx.step(0.01, 100); ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(f"flags={flags}: best frame {t*1e3:.2f} ms / 100 steps = {t*1e4:.1f} us per step -> {V*100/t/1e9:.1f} G/s")
    ctx.close()
