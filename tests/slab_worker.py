"""One rank of a 2-process slab ring (tests/test_gpu_slab.py::test_two_process_ring_over_ipc):
    slab_worker.py <session> <rank> <nranks> <dims,comma> <out.npz>
Device = rank.  Same inputs as the parent's oracle run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import stochquant_b200 as sq  # noqa: E402
from slab_common import run_rank  # noqa: E402

name, rank, nranks, dims, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
dims = tuple(int(d) for d in dims.split(","))
rng = np.random.default_rng(9)
phi0 = rng.normal(size=int(np.prod(dims))) * 0.5
r = run_rank(sq, name, rank, nranks, dims, phi0, [5, 20], 0.01, device=rank, real="f32", math="accurate", pot=4,
             m2=0.25, lam=0.5)
r.pop("stats")
np.savez(out, **r)
