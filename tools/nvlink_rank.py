"""One rank of a two-process time-slab ring for the NVLink byte-counter capture (tools/nvlink_capture.sh):
    nvlink_rank.py <session> <rank> <nranks>
256 x 256 x 64 slices (16.8 MB each), 8 per rank, 3 tau-steps; device = rank.  Prints the ring's own accounting."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochquant_b200 as sq  # noqa: E402
from stochquant_b200.slabs import run_rank  # noqa: E402

name, rank, nranks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dims = (256, 256, 64, 8 * nranks)
r = run_rank(sq, name, rank, nranks, dims, None, [3], 0.01, device=rank, real="f32", math="fast")
vs = dims[0] * dims[1] * dims[2]
print(f"rank {rank}: seed {r['seed']} events {r['nevents']} halo slice {vs * 4} bytes, two faces per tau-step = {2 * vs * 4} bytes sent; stats {r['stats']}")
