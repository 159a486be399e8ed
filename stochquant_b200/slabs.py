"""Host-side helpers of the time-slab decomposition: how a lattice's time extent is cut
over the ranks of a ring, and one rank's life cycle through the C-ABI (session -> context -> join ->
frames -> measure).  Used by bench.py's ring workloads and by the slab tests."""
from __future__ import annotations

import numpy as np


def split_slabs(Lt: int, nranks: int):
    """Time slices per rank, as even as possible, in rank order: [(t0, nt), ...]."""
    base, extra = divmod(Lt, nranks)
    out, t = [], 0
    for r in range(nranks):
        nt = base + (1 if r < extra else 0)
        out.append((t, nt))
        t += nt
    return out


def run_rank(sq, name, rank, nranks, dims, phi0, frames, dtau, device=0, real="f32", math="accurate", pot=0,
             m2=0.0, lam=0.0, seed=1242608872, flags=0):
    """One rank of the ring: frames = list of tau-step counts (one sq_step each).  Returns a dict of
    numpy arrays / ints for this rank's slab."""
    t0, nt = split_slabs(dims[-1], nranks)[rank]
    vs = int(np.prod(dims[:-1]))
    sess = sq.Session(name, rank, nranks)
    ctx = sq.Context(dims, real=real, math=math, potential=pot, m2=m2, lam=lam, seed=seed, device=device,
                     slab=(t0, nt), flags=flags)
    try:
        if phi0 is not None:
            ctx.upload(np.asarray(phi0, dtype=ctx.dtype).reshape(-1)[t0 * vs:(t0 + nt) * vs])
        ctx.join(sess)
        for n in frames:
            assert ctx.step(dtau, n)
        m = ctx.measure()
        out = {"t0": t0, "nt": nt, "field": ctx.download(), "seed": int(m["seed"]), "runs": int(m["runs"]),
               "slice_x": m["slice_x"].copy(), "slice_xx0": m["slice_xx0"].copy(), "corr": m["corr"].copy(),
               "nevents": int(m["nevents"]), "nclamped": int(m["nclamped"]), "mean_phi": m["mean_phi"],
               "mean_phi2": m["mean_phi2"], "stats": ctx.slab_stats()}
    except Exception:
        sess.abort()
        raise
    finally:
        ctx.close()
        sess.close()
    return out
