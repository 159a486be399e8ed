"""tau-step time of a streaming workload for the SQ_FIN_BATCH / SQ_PDL in the environment, plus a digest of everything the
run leaves behind (field, seed, running means) -- the digests of different settings must be equal (the grouping of the
finalize launches only moves work between streams).  Usage: SQ_FIN_BATCH=4 python tools/exp_fin_batch.py c3|slab|c5"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stochquant_b200 as sq

w = sys.argv[1] if len(sys.argv) > 1 else "c3"
kw, loops = {"c3": (dict(dims=(64, 64, 64, 64)), 100), "slab": (dict(dims=(256, 256, 256, 32)), 10),
             "c5": (dict(dims=(32, 32, 32, 32), nchains=64, potential=4, m2=0.25, lam=0.5), 20),
             "small": (dict(dims=(32, 16, 8, 16)), 37)}[w]
ctx = sq.Context(real="f32", math="fast", **kw)
V = int(np.prod(kw["dims"])) * kw.get("nchains", 1)
h = hashlib.sha256()
for n in (37, 5, 1, 8):  # odd sequence lengths: partial groups, the join at the end of a sequence
    ctx.step(0.01, n)
    m = ctx.measure()
    f = ctx.download()
    if f.nbytes > (64 << 20):  # big fields: an order-independent 64-bit sum and xor of the words instead of sha256
        u = f.view(np.uint32)
        h.update(repr((int(u.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(u)))).encode())
    else:
        h.update(f.tobytes())
    for k in ("slice_x", "slice_xx0", "corr"):
        h.update(np.ascontiguousarray(m[k]).tobytes())
    h.update(repr((m["seed"], m["mean_phi"], m["mean_phi2"], m["runs"], m["nclamped"])).encode())
ts = []
for _ in range(8):
    t0 = time.perf_counter(); ctx.step(0.01, loops); ts.append(time.perf_counter() - t0)
dt = float(np.median(ts))
print(f"{w} SQ_FIN_BATCH={os.environ.get('SQ_FIN_BATCH')} SQ_PDL={os.environ.get('SQ_PDL')}: {dt / loops * 1e6:.2f} us per tau-step, "
      f"{V * loops / dt / 1e9:.1f} G site-updates/s, events {ctx.measure()['nevents']}, digest {h.hexdigest()[:16]}")
ctx.close()
