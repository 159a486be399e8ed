"""GPU parity of the multi-GPU slab decomposition (north_star (4)): a ring of contexts,
each owning a slab of time slices, must reproduce the whole-lattice oracle (oracle/sq_oracle.c:
sqo_lattice_step) -- seeds bit-exact, fields within the single-GPU tolerances, running means of the
slice observables across slabs.

Rings of 1-3 ranks run as host threads on ONE GPU (plain peer pointers inside the process: the halo
protocol, the finder and the event agreement are the same code as across GPUs); the 2-process ring
over CUDA IPC needs two GPUs and runs where they exist (gpurun --gpus 2)."""
import os
import subprocess
import sys
import threading
import uuid

import numpy as np
import pytest

from helpers import maxabs, seed_with_retry_at
from slab_common import run_rank, split_slabs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ATOL = {("f64", "accurate"): 5e-7, ("f32", "accurate"): 2e-5, ("f32", "fast"): 6e-5}
DTAU = 0.01


def ring_threads(sq, nranks, dims, phi0, frames, **kw):
    name = "t" + uuid.uuid4().hex[:12]
    res, err = [None] * nranks, []

    def work(r):
        try:
            res[r] = run_rank(sq, name, r, nranks, dims, phi0, frames, DTAU, **kw)
        except Exception as e:  # noqa: BLE001
            err.append((r, e))
    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not err, err
    assert all(r is not None for r in res)
    return res


def check_against_oracle(res, o, dims, tol_field, tol_obs):
    vs = int(np.prod(dims[:-1]))
    tm = dims[-1] // 2
    for r in res:
        t0, nt = r["t0"], r["nt"]
        assert r["seed"] == o.seed, "step seed must be bit-exact on every rank"
        assert r["runs"] == o.L.runs
        err = maxabs(r["field"], o.field[t0 * vs:(t0 + nt) * vs])
        assert err < tol_field, (t0, err)
        assert maxabs(r["slice_x"], o.slice_x[t0:t0 + nt]) < tol_obs
        assert maxabs(r["slice_xx0"], o.slice_xx0[t0:t0 + nt]) < tol_obs
        assert maxabs(r["corr"], (o.slice_xx0 - o.slice_x * o.slice_x[tm])[t0:t0 + nt]) < tol_obs
        assert r["nclamped"] == 0


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("dims,real,math,pot", [((16, 8, 4, 12), "f32", "accurate", 4), ((64, 24), "f64", "accurate", 0),
                                               ((32, 8, 9), "f32", "fast", 4)])
def test_ring_vs_oracle(gpu_sq, oracle, nranks, dims, real, math, pot):
    rng = np.random.default_rng(5)
    phi0 = rng.normal(size=int(np.prod(dims))) * 0.5
    o = oracle.LatticeOracle(dims, real=oracle.F32 if real == "f32" else oracle.F64, potential=pot, m2=0.25, lam=0.5,
                             phi0=phi0)
    frames = [1, 2, 14]
    res = ring_threads(gpu_sq, nranks, dims, phi0, frames, real=real, math=math, pot=pot, m2=0.25, lam=0.5)
    o.step(DTAU, sum(frames))
    check_against_oracle(res, o, dims, ATOL[(real, math)], 50 * ATOL[(real, math)])
    # one finder scan per step and rank at least
    assert all(r["stats"]["finder_scans"] >= sum(frames) for r in res)


@pytest.mark.parametrize("where", ["slab0", "slab1", "slab2_last", "omega", "plus"])
def test_ring_events_agreed(gpu_sq, oracle, where):
    """An RNG event (inf-retry / `seed+=`) in one slab changes the seeds of every later draw on
    every rank: found by the finder, agreed through the session, applied everywhere."""
    dims = (16, 4, 12)
    V = 16 * 4 * 12
    vs = 64
    seed = {"plus": 39512}.get(where)
    if seed is None:
        g = {"slab0": 3 * vs + 5, "slab1": 5 * vs + 17, "slab2_last": V - 1, "omega": V}[where]
        seed = seed_with_retry_at(oracle, g)
    rng = np.random.default_rng(6)
    phi0 = rng.normal(size=V) * 0.5
    o = oracle.LatticeOracle(dims, real=oracle.F64, potential=0, seed=seed, phi0=phi0)
    res = ring_threads(gpu_sq, 3, dims, phi0, [6], real="f64", math="accurate", seed=seed)
    o.step(DTAU, 6)
    assert o.L.nevents >= 1
    check_against_oracle(res, o, dims, ATOL[("f64", "accurate")], 1e-5)
    assert all(r["nevents"] >= 1 for r in res)


def test_ring_equals_single_context(gpu_sq):
    """Same lattice, whole vs 4 slabs: the fields must be IDENTICAL bit for bit (same kernel, same
    operation order, same stream) and so must the slice sums."""
    dims = (32, 8, 4, 16)
    rng = np.random.default_rng(7)
    phi0 = (rng.normal(size=int(np.prod(dims))) * 0.5).astype(np.float32)
    whole = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5)
    whole.upload(phi0)
    whole.step(DTAU, 11)
    ref = whole.download()
    res = ring_threads(gpu_sq, 4, dims, phi0, [11], real="f32", math="fast", pot=4, m2=0.25, lam=0.5)
    got = np.concatenate([r["field"] for r in res])
    assert np.array_equal(got, ref)
    m = whole.measure()
    assert maxabs(np.concatenate([r["slice_x"] for r in res]), m["slice_x"]) < 1e-12
    assert maxabs(np.concatenate([r["slice_xx0"] for r in res]), m["slice_xx0"]) < 1e-12


@pytest.mark.parametrize("dims,nranks", [((32, 32, 4, 12), 3), ((64, 16, 16, 8), 2), ((256, 8, 4, 8), 4)])
def test_ring_rowblock_kernel_equals_single_context(gpu_sq, oracle, dims, nranks):
    """The slab ring on the row-block staging kernel (SQ_FLAG_ROWBLOCK_KERNEL: the producer warp waits for the
    neighbours' flags, raises them when a boundary slice's last tile retires; the computing warps push the
    boundary slices): bit-identical to one context on the default tile kernel, with an RNG event on the way."""
    V = int(np.prod(dims))
    rng = np.random.default_rng(8)
    phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
    seed = seed_with_retry_at(oracle, V // 3 + 5)
    whole = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, seed=seed)
    whole.upload(phi0)
    whole.step(DTAU, 5)
    whole.step(DTAU, 6)
    ref, seed_ref = whole.download(), whole.measure()["seed"]
    whole.close()
    res = ring_threads(gpu_sq, nranks, dims, phi0, [5, 6], real="f32", math="fast", pot=4, m2=0.25, lam=0.5, seed=seed,
                       flags=gpu_sq.SQ_FLAG_ROWBLOCK_KERNEL)
    assert all(r["seed"] == seed_ref for r in res)
    assert np.array_equal(np.concatenate([r["field"] for r in res]), ref)
    assert sum(r["nevents"] for r in res) >= nranks  # every rank replayed the event


def test_ring_long_run_equals_single_context_and_oracle(gpu_sq, oracle):
    """Error growth over a long run on a ring (nothing beyond 25 steps): 3 slabs, 60 + 140 tau-steps in two
    frames, still IDENTICAL bit for bit to one context -- and that context stays within the per-step tolerance times a
    modest growth factor of the fp64-noise oracle, seeds bit-exact after 200 steps."""
    dims = (32, 8, 4, 12)
    rng = np.random.default_rng(9)
    phi0 = (rng.normal(size=int(np.prod(dims))) * 0.5).astype(np.float32)
    kw = dict(real="f32", math="accurate", pot=4, m2=0.25, lam=0.5)
    whole = gpu_sq.Context(dims, real="f32", math="accurate", potential=4, m2=0.25, lam=0.5)
    whole.upload(phi0)
    whole.step(DTAU, 60)
    whole.step(DTAU, 140)
    ref, mref = whole.download(), whole.measure()
    whole.close()
    res = ring_threads(gpu_sq, 3, dims, phi0, [60, 140], **kw)
    assert np.array_equal(np.concatenate([r["field"] for r in res]), ref)
    assert all(r["seed"] == mref["seed"] for r in res)
    o = oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=0.5, phi0=phi0)
    o.step(DTAU, 200)
    assert o.seed == mref["seed"]
    assert maxabs(ref, o.field) < 20 * ATOL[("f32", "accurate")]


def test_multi_process_ring_over_ipc(gpu_sq, oracle, tmp_path):
    """One process per GPU (all GPUs of the box, up to 8), halo arenas mapped through CUDA IPC, NVLink
    peer stores: needs >= 2 GPUs."""
    ngpu = gpu_sq.load().sq_device_count()
    if ngpu < 2:
        pytest.skip("needs two or more GPUs (gpurun --gpus 2|4|8)")
    nr = min(ngpu, 8)
    dims = (32, 16, 8, 24)
    name = "p" + uuid.uuid4().hex[:12]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "slab_worker.py"), name, str(r), str(nr),
                               ",".join(map(str, dims)), str(tmp_path / f"r{r}.npz")]) for r in range(nr)]
    for p in procs:
        assert p.wait(300) == 0
    rng = np.random.default_rng(9)
    phi0 = rng.normal(size=int(np.prod(dims))) * 0.5
    o = oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=0.5, phi0=phi0)
    o.step(DTAU, 25)
    res = []
    for r, (t0, nt) in enumerate(split_slabs(dims[-1], nr)):
        z = np.load(tmp_path / f"r{r}.npz")
        res.append({k: (z[k] if z[k].ndim else z[k].item()) for k in z.files} | {"t0": t0, "nt": nt})
    check_against_oracle(res, o, dims, ATOL[("f32", "accurate")], 50 * ATOL[("f32", "accurate")])


def test_c4_slab_size_ring_equals_single_context(gpu_sq):
    """configs[3] geometry at the per-GPU slab size of the 8-GPU run, halved in time to keep the test
    short: 256^3 slices (67 MB each), 16 of them, as a joined ring of ONE rank (session, halo pushes into
    its own arena, finder) against the plain single context: bit-identical fields and seed, although RNG
    events (0.08 expected per step at this volume) are handled by two different mechanisms -- replay
    there, finder here.  (Several ranks as threads on ONE GPU are only safe for small lattices: a big
    rank's spinning boundary CTAs can fill every SM slot and starve the neighbour it waits for; with
    one rank per GPU that cannot happen -- test_multi_process_ring_over_ipc covers real rings.)"""
    dims = (256, 256, 256, 16)
    whole = gpu_sq.Context(dims, real="f32", math="fast")
    whole.step(DTAU, 6)
    ref = whole.download()
    seed = whole.measure()["seed"]
    whole.close()
    res = ring_threads(gpu_sq, 1, dims, None, [6], real="f32", math="fast")
    assert res[0]["seed"] == seed
    assert np.array_equal(res[0]["field"], ref)


@pytest.mark.parametrize("gid", [0, 3, 4, 2047, 2048, 8191, 8192, 16383, 16384])
def test_ring_events_on_marching_kernel(gpu_sq, oracle, gid):
    """Forced events on a marching-kernel shape in a ring of two (slab boundary at gid 8192): found by the
    finder, agreed through the session, applied through virtual start seeds on both ranks."""
    dims = (32, 8, 8, 8)
    V = int(np.prod(dims))
    seed = seed_with_retry_at(oracle, gid)
    rng = np.random.default_rng(22)
    phi0 = rng.normal(size=V) * 0.5
    o = oracle.LatticeOracle(dims, real=oracle.F32, potential=0, seed=seed, phi0=phi0)
    res = ring_threads(gpu_sq, 2, dims, phi0, [4], real="f32", math="fast", seed=seed)
    o.step(DTAU, 4)
    assert o.L.nevents >= 1
    check_against_oracle(res, o, dims, ATOL[("f32", "fast")], 50 * ATOL[("f32", "fast")])
    assert all(r["nevents"] >= 1 for r in res)


@pytest.mark.parametrize("nranks", [2, 4])
def test_ring_thin_slabs(gpu_sq, oracle, nranks):
    """Four time slices over 2 or 4 ranks: with one slice per rank the same slice is the lower AND the upper
    boundary (it waits on both neighbours and pushes to both); with two there is no interior at all."""
    dims = (32, 8, 8, 4)
    rng = np.random.default_rng(23)
    phi0 = rng.normal(size=int(np.prod(dims))) * 0.5
    o = oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=0.5, phi0=phi0)
    res = ring_threads(gpu_sq, nranks, dims, phi0, [3, 8], real="f32", math="fast", pot=4, m2=0.25, lam=0.5)
    o.step(DTAU, 11)
    check_against_oracle(res, o, dims, ATOL[("f32", "fast")], 50 * ATOL[("f32", "fast")])
