"""GPU parity on the configurations bench.py TIMES, end to end, and on runs in which the clamp fires.

The bench's c2 step is ONE 1000-step sq_step on 1024^2 (a single resident launch with checkpoints and the
closed-form running means), c3 a 100-step sequence on 64^4: exactly those calls are compared here with the
OpenMP oracle (oracle/sq_oracle.c: sqo_lattice_step_omp; definition in DESIGN.md section 4) -- step seed
bit-exact, field within the per-mode tolerance below, slice observables and the correlator.

Clamp (tau_kernel.cl:122-132): with a huge noise amplitude and a sub-critical step size the clamp fires at
~16 % of the sites every step while the drift stays contractive, so GPU and oracle can be compared site by
site: ACCURATE must agree on every clamp decision (nclamped equal); FAST differs from the oracle by the SFU's
error on the draw (relative 4e-4 of the noise amplitude in the worst case), which can flip a decision only
for a value within that distance of +-1000."""
import numpy as np
import pytest

from helpers import maxabs, seed_with_retry_at

pytestmark = pytest.mark.gpu

DTAU = 0.01
# max over all sites / rms, after 10^3 steps (the drift is contractive: errors do not accumulate)
TOL_MAX = {"fast": 6e-5, "accurate": 2e-5}
TOL_RMS = {"fast": 2e-6, "accurate": 5e-7}


def _compare(g, o, math, obs_tol=2e-4):
    m = g.measure()
    assert m["seed"] == o.seed, "step seed must be bit-exact"
    assert m["runs"] == o.L.runs
    d = np.abs(g.download().astype(np.float64) - o.field.astype(np.float64))
    assert d.max() < TOL_MAX[math], d.max()
    assert np.sqrt(np.mean(d ** 2)) < TOL_RMS[math]
    tm = g.dims[-1] // 2
    assert maxabs(m["slice_x"], o.slice_x) < obs_tol
    assert maxabs(m["slice_xx0"], o.slice_xx0) < obs_tol
    assert maxabs(m["corr"], o.slice_xx0 - o.slice_x * o.slice_x[tm]) < obs_tol
    fld = o.field.astype(np.float64)
    assert abs(m["mean_phi"] - fld.mean()) < obs_tol and abs(m["mean_phi2"] - (fld ** 2).mean()) < obs_tol
    assert m["nclamped"] == 0
    return m


def test_c2_bench_step_vs_oracle(gpu_sq, oracle):
    """configs[1] exactly as bench.py runs it: 1024^2 fp32, cold start, seed 1242608872, dtau 0.01, ONE
    1000-step sq_step, then a second one (resume from the device state), FAST and ACCURATE."""
    dims = (1024, 1024)
    o = oracle.LatticeOracle(dims, real=oracle.F32, seed=1242608872)
    gs = {m: gpu_sq.Context(dims, real="f32", math=m, seed=1242608872) for m in ("fast", "accurate")}
    for frame in range(2):
        o.step(DTAU, 1000, omp=True)
        for math, g in gs.items():
            assert g.step(DTAU, 1000)
            m = _compare(g, o, math)
            assert m["runs"] == 1000 * (frame + 1)
            assert m["nevents"] == o.L.nevents  # the same chain events met and replayed
    for g in gs.values():
        g.close()


def test_c2_bench_step_with_event_mid_launch(gpu_sq, oracle):
    """The same 1000-step launch with an inf-retry forced into step 700 at a site in the middle of the lattice:
    the launch is left early, resumed from the checkpoint of step 640, the event taken by a streaming step."""
    from test_gpu_lattice import _seed_with_event_in_step
    dims = (1024, 1024)
    V = 1024 * 1024
    S = _seed_with_event_in_step(oracle, V, 524288 + 77, 700)
    o = oracle.LatticeOracle(dims, real=oracle.F32, seed=S)
    g = gpu_sq.Context(dims, real="f32", math="fast", seed=S)
    o.step(DTAU, 1000, omp=True)
    assert g.step(DTAU, 1000)
    m = _compare(g, o, "fast")
    assert o.L.nevents >= 1 and m["nevents"] == o.L.nevents


@pytest.mark.parametrize("math", ["fast", "accurate"])
def test_c3_bench_step_vs_oracle(gpu_sq, oracle, math):
    """configs[2] as bench.py runs it: 64^4 fp32, cold start, one 100-step sq_step."""
    dims = (64, 64, 64, 64)
    o = oracle.LatticeOracle(dims, real=oracle.F32, seed=1242608872)
    g = gpu_sq.Context(dims, real="f32", math=math, seed=1242608872)
    o.step(DTAU, 100, omp=True)
    assert g.step(DTAU, 100)
    m = _compare(g, o, math)
    assert m["nevents"] == o.L.nevents


# --------------------------------------------------------------------------------------------------------
# clamp parity
CLAMP_C = 5000.0  # noise amplitude C sqrt(2 dtau) = 707: |v| > 1000 at ~16 % of the sites per step
CLAMP_SHAPES = [((32, 32), "f32"), ((32, 32), "f64"), ((128, 40), "f32"), ((32, 8, 8, 8), "f32"), ((16, 12, 10), "f64")]


@pytest.mark.parametrize("math", ["accurate", "fast"])
@pytest.mark.parametrize("dims,real", CLAMP_SHAPES)
def test_clamp_fires_and_matches_oracle(gpu_sq, oracle, dims, real, math):
    """Generic (2-D / fp64), resident (128x40) and marching (32x8x8x8) kernels with the clamp firing."""
    g = gpu_sq.Context(dims, real=real, math=math, potential=0, noise_c=CLAMP_C, seed=1242608872)
    o = oracle.LatticeOracle(dims, real=oracle.F32 if real == "f32" else oracle.F64, potential=0, c=CLAMP_C,
                             seed=1242608872)
    V = int(np.prod(dims))
    for n in (1, 4, 20):
        assert g.step(DTAU, n)
        o.step(DTAU, n)
        m = g.measure()
        assert m["seed"] == o.seed
        a, b = g.download().astype(np.float64), o.field.astype(np.float64)
        assert np.all(np.abs(a) <= 1000.0)
        assert o.L.nclamped > 0.05 * V * o.L.runs
        d = np.abs(a - b)
        if math == "accurate":
            # cosf/logf differ by <= 2 ulp of fp32 between CUDA and glibc: 1e-7 of an amplitude of 707 (x5 sigma)
            assert d.max() < 2e-3, d.max()
            assert np.array_equal(np.abs(a) == 1000.0, np.abs(b) == 1000.0)
            assert m["nclamped"] == o.L.nclamped
        else:
            flipped = (np.abs(a) == 1000.0) != (np.abs(b) == 1000.0)
            assert flipped.sum() <= 2
            assert d[~flipped].max() < 1.0 and np.sqrt(np.mean(d[~flipped] ** 2)) < 2e-2
            assert abs(m["nclamped"] - o.L.nclamped) <= 4
        assert maxabs(m["slice_x"], o.slice_x) < 1e-2


def test_clamp_count_survives_event_replay(gpu_sq, oracle):
    """A step that is replayed after an RNG event must not count its clamp hits twice (streaming kernels:
    per-step clamp slots, committed only for steps that stand)."""
    dims = (32, 8, 8, 8)
    seed = seed_with_retry_at(oracle, 4099)
    g = gpu_sq.Context(dims, real="f32", math="accurate", potential=0, noise_c=CLAMP_C, seed=seed)
    o = oracle.LatticeOracle(dims, real=oracle.F32, potential=0, c=CLAMP_C, seed=seed)
    g.step(DTAU, 6)
    o.step(DTAU, 6)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1 and m["seed"] == o.seed
    assert m["nclamped"] == o.L.nclamped and o.L.nclamped > 0


def test_compat1d_clamp_fires_and_matches_oracle(gpu_sq, oracle):
    """The reference kernel's own clamp (tau_kernel.cl:122-132) in the 1-D run.  C = 20000 makes
    |dw| ~ 2000 |r|, so most sites are clamped in the very first step.
    (a) potID 0, N = 17 / 100: the first one-step frame is ACCEPTED with clamped sites in it -- the committed
        field must equal the oracle's site by site, +-1000 at the same sites; the next frame trips the
        stability test (:135-143) and is rolled back.
    (b) potID 3, N = 200 (the default run's shape): rejected frames; seed, lrgEl, lrgVl (which keep their
        values across a rollback, tauhost.c:533-554) and the executed step count must agree."""
    c = 20000.0
    for N, dt, dtau, pot in ((17, .05, 5e-4, 0), (100, .1, 3e-3, 0), (200, .02, 1e-4, 3)):
        f, om, r1 = oracle.host_init(N, dt, dtau)
        g = gpu_sq.Context([N], kernel="compat1d", potential=pot, spacing=dt, noise_c=c, f0=f, omega0=om, seed=r1)
        o = oracle.Compat1D(N, dt, dtau, pot, c, f, om, r1)
        accepted_clamped = 0
        for loops in (1, 30):
            st_o = o.frame(loops)
            st_g = g.step(dtau, loops)
            assert st_g == st_o
            m = g.measure()
            assert m["seed"] == o.s.rand1 and m["lrgEl"] == o.s.lrgEl
            assert abs(m["lrgVl"] - o.s.lrgVl) < 1e-9 * max(1.0, abs(o.s.lrgVl))
            # |dw| up to ~1e4 before the clamp: the fp32 transcendentals' 1e-7 becomes ~1e-3 absolute
            assert maxabs(m["f"], o.f) < 5e-3 and maxabs(m["x"], o.x) < 5e-3 and maxabs(m["xx0"], o.xx0) < 5.0
            assert np.array_equal(np.abs(m["f"]) == 1000.0, np.abs(o.f) == 1000.0)
            if st_o:
                accepted_clamped += int((np.abs(o.f) == 1000.0).sum())
        if pot == 0:
            assert accepted_clamped > 0
        g.close()


def test_streaming_event_in_later_step_keeps_every_sample(gpu_sq, oracle):
    """Streaming path (4-D marching kernel and the generic 2-D kernel via SQ_FLAG_FORCE_STREAMING): an event in
    step k > 0 while finalize(k-1) may still be running on the side stream.  Every step before the event step
    must contribute its sample to the running means (finalize skips only the event step and later ones)."""
    for dims, flags, gid in (((32, 8, 8, 8), 0, 4099), ((64, 32), 2, 777)):
        V = int(np.prod(dims))
        from test_gpu_lattice import _seed_with_event_in_step
        for step in (1, 2, 5):
            S = _seed_with_event_in_step(oracle, V, gid, step)
            rng = np.random.default_rng(5)
            phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
            g = gpu_sq.Context(dims, real="f32", math="accurate", seed=S, flags=flags)
            o = oracle.LatticeOracle(dims, real=oracle.F32, seed=S, phi0=phi0)
            g.upload(phi0)
            g.step(DTAU, 9)
            o.step(DTAU, 9)
            m = g.measure()
            assert o.L.nevents >= 1 and m["nevents"] >= 1 and m["seed"] == o.seed and m["runs"] == 9
            assert maxabs(m["slice_x"], o.slice_x) < 1e-6 and maxabs(m["slice_xx0"], o.slice_xx0) < 1e-6
            g.close()
