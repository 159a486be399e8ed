"""Experiment behind DESIGN.md section 8 (no temporal blocking for d >= 3): the ceiling of ANY scheme that turns HBM traffic into
on-chip hits is the streaming kernel's speed on a lattice that already sits in L2.  Kernel time per site-update (events
around the update launches, sq_kernel_timing) for 4-D lattices from L2-resident to HBM-streaming."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stochquant_b200 as sq
for dims in ((64, 64, 64, 8), (64, 64, 64, 16), (64, 64, 64, 32), (64, 64, 64, 64), (256, 256, 64, 8), (256, 256, 256, 8), (256, 256, 256, 32)):
    V = int(np.prod(dims))
    ctx = sq.Context(dims, real="f32", math="fast")
    n = 40 if V < 1e8 else 10
    for _ in range(2):
        ctx.step(0.01, n)
    ctx.kernel_timing(True)
    for _ in range(3):
        ctx.step(0.01, n)
    kms, kn = ctx.kernel_time()
    ctx.kernel_timing(False)
    us = 1e3 * kms / kn
    print(f"{'x'.join(map(str, dims)):>16s}: {2 * V * 4 / 2**20:8.0f} MiB in two buffers  {us:9.1f} us/step  {1e6 * us / V:6.2f} ps/site  {V / us / 1e3:6.1f} G site-updates/s  ({kn} launches)")
    ctx.close()
