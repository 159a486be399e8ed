"""Experiment: frame time with and without per-launch event timing (sq_kernel_timing)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochquant_b200 as sq
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256,256,256,32").split(","))
ctx = sq.Context(dims, real="f32", math="fast")
import numpy as np
V = int(np.prod(dims))
for _ in range(3): ctx.step(0.01, 10)
for timing in (False, True, False, True):
    ctx.kernel_timing(timing)
    t0 = time.perf_counter(); ctx.step(0.01, 10); dt = time.perf_counter() - t0
    kms, kn = ctx.kernel_time()
    print(f"timing={timing}: frame {dt*1e3:.2f} ms for 10 steps -> {V*10/dt/1e9:.1f} G/s; kernel events: {kms:.2f} ms over {kn} launches; nevents {ctx.measure()['nevents']}")
