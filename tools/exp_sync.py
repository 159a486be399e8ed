"""Experiment: where does the host time of one frame go? (c2: 1000 steps per frame)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochquant_b200 as sq, numpy as np, torch
dims = (1024, 1024)
ctx = sq.Context(dims, real="f32", math="fast")
V = int(np.prod(dims)); n = 1000
for _ in range(3): ctx.step(0.01, n)
stream = torch.cuda.ExternalStream(ctx.stream)
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter(); e0.record(stream); ctx.step_async(0.01, n); t1 = time.perf_counter(); e1.record(stream)
    ctx.sync(); t2 = time.perf_counter(); e2.record(stream); e2.synchronize(); t3 = time.perf_counter()
    print(f"host: async {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms, tail {1e3*(t3-t2):.3f}; gpu: e0-e1 {e0.elapsed_time(e1):.3f} ms, e0-e2 {e0.elapsed_time(e2):.3f} ms")
