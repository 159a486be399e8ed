#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE kernel itself.

Runs the reference's own source (/root/reference/tau_kernel.cl, compiled as C into
oracle/_ref/libtau_ref.so by oracle/Makefile) under the serial work-item schedule and stores
its outputs as small fixtures.  /root/reference does not exist on the GPU box, so the tests
read these files instead; re-run this script in the dev container to regenerate.

  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def random_kats():
    R = O.ref()
    rng = np.random.default_rng(20261018)
    seeds, gids, vals, after = [], [], [], []
    special = [(1242608872, 0), (1242608872, 1), (1242608872, 200), (0, 0), (1, 0), (2**64 - 5, 7),
               (1760221443, 3), (668289095, 3), (177446488061229, 0), (177446488061224, 5)]
    cases = special + [(int(rng.integers(0, 2**48)), int(rng.integers(0, 2**12))) for _ in range(2000)]
    cases += [(int(rng.integers(0, 2**31)), int(rng.integers(0, 256))) for _ in range(500)]
    for s, g in cases:
        a = C.c_ulong(s)
        v = R.sq_ref_random(C.byref(a), g)
        seeds.append(s); gids.append(g); vals.append(v); after.append(a.value)
    np.savez_compressed(os.path.join(HERE, "random_kat.npz"), seed=np.array(seeds, dtype=np.uint64),
                        gid=np.array(gids, dtype=np.int64), value=np.array(vals),
                        seed_after=np.array(after, dtype=np.uint64))


def kernel_runs():
    out = {}
    cases = {"dw200": (3, 200, .02, 1e-4, 60), "ho100": (0, 100, .1, 3e-3, 60), "dw17": (3, 17, .05, 5e-4, 40),
             "dw200_unstable": (3, 200, .02, 2e-3, 8)}
    for name, (pot, N, dt, dtau, steps) in cases.items():
        f, om, r1 = O.host_init(N, dt, dtau)
        r = O.RefKernel(N, dt, dtau, pot, 1.0, f, om, r1)
        seeds = []
        ok = True
        for _ in range(steps):
            ok = r.steps_canonical(1)
            seeds.append(r.rand1.value)
            if not ok:
                break
        out[name + "_params"] = np.array([pot, N, dt, dtau, steps], dtype=np.float64)
        out[name + "_f0"] = f
        out[name + "_omega0"] = np.array([om])
        out[name + "_seed0"] = np.array([r1], dtype=np.uint64)
        out[name + "_newf"] = r.newf.copy()
        out[name + "_newx"] = r.newx.copy()
        out[name + "_newxx0"] = r.newxx0.copy()
        out[name + "_omega"] = np.array([r.omega.value])
        out[name + "_seeds"] = np.array(seeds, dtype=np.uint64)
        out[name + "_lrg"] = np.array([r.lrgEl.value, r.lrgVl.value, r.stable.value], dtype=np.float64)
        # {chain, in-place}: one launch with Loops=steps (what a CPU OpenCL runtime does)
        r2 = O.RefKernel(N, dt, dtau, pot, 1.0, f, om, r1)
        r2.launch(steps)
        out[name + "_inplace_newf"] = r2.newf.copy()
        out[name + "_inplace_newx"] = r2.newx.copy()
        out[name + "_inplace_seed"] = np.array([r2.rand1.value], dtype=np.uint64)
        out[name + "_inplace_lrg"] = np.array([r2.lrgEl.value, r2.lrgVl.value, r2.stable.value], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "time_dev_ref.npz"), **out)


def model_fns():
    R = O.ref()
    rng = np.random.default_rng(7)
    a = rng.uniform(-3, 8, 400)
    w = rng.uniform(0, 4, 400)
    cl = np.array([R.clas(x, y, 3) for x, y in zip(a, w)])
    dd = np.array([R.ddPot(x, 3) for x in cl])
    np.savez_compressed(os.path.join(HERE, "model_fns.npz"), a=a, w=w, clas3=cl, ddpot3=dd,
                        intconst=np.array([R.intConst(0), R.intConst(3)]))


if __name__ == "__main__":
    if not O.ref_available():
        sys.exit("oracle/_ref missing: needs /root/reference (dev container)")
    random_kats()
    kernel_runs()
    model_fns()
    print("golden fixtures written to", HERE)
