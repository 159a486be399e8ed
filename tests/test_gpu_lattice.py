"""GPU parity: the d-dimensional lattice kernels (SQ_KERNEL_LATTICE) through the C-ABI against the
oracle's definition (oracle/sq_oracle.c: sqo_lattice_step).

Bit-exact: the LCG stream per site, the step seeds, event replay.  Tolerance: field values --
ACCURATE differs from the oracle only by CUDA-vs-glibc cosf/logf (<=2 ulp fp32 ~ 1e-7 on r);
FAST (SFU log2/cos) by <= ~2e-6 on r.  With noise amplitude sqrt(2 dtau) ~ 0.14 and a
contractive drift the per-site error stays below the ATOLs here for the step counts used."""
import numpy as np
import pytest

from helpers import maxabs, seed_with_retry_at

pytestmark = pytest.mark.gpu

# max over ALL sites; FAST's bound is set by rare draws with v1 within 2^-12 of 1 (sq_noise.cuh)
ATOL = {("f64", "accurate"): 5e-7, ("f32", "accurate"): 2e-5, ("f64", "fast"): 5e-5, ("f32", "fast"): 6e-5}
DTAU = 0.01


def pair(gpu_sq, oracle, dims, real="f32", math="accurate", pot=0, m2=0.0, lam=0.0, seed=1242608872, phi0=None, **kw):
    g = gpu_sq.Context(dims, real=real, math=math, potential=pot, m2=m2, lam=lam, seed=seed, **kw)
    o = oracle.LatticeOracle(dims, real=oracle.F32 if real == "f32" else oracle.F64, potential=pot, m2=m2, lam=lam,
                             seed=seed, phi0=phi0)
    if phi0 is not None:
        g.upload(np.asarray(phi0, dtype=g.dtype))
    return g, o


def test_site_stream_bit_exact(gpu_sq, oracle):
    """(t1,t2) per site from the device jump-ahead == literal chain in gid order (tau_kernel.cl:273-281)."""
    g, _ = pair(gpu_sq, oracle, (64, 64))
    t1, t2 = g.debug_draws(0, 4096)
    o1, o2, _ = oracle.lattice_draws(1242608872, 4096)
    assert np.array_equal(t1, o1) and np.array_equal(t2, o2)
    # far into a 256^4-sized lattice: gid0 = 2^32 - 50 crosses the 32-bit boundary
    g0 = 2**32 - 50
    t1, t2 = g.debug_draws(g0, 100)
    s = oracle.lib().sqo_jump(1242608872, 0, g0)
    for k in range(100):
        _, rec = oracle.random(s, g0 + k)
        assert (int(t1[k]), int(t2[k])) == (rec.t1, rec.t2)
        s = rec.seed_after


@pytest.mark.parametrize("math", ["accurate", "fast"])
@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("dims,pot", [((64, 48), 0), ((16, 12, 10), 4), ((8, 6, 4, 10), 4), ((12, 10), 4),
                                      ((32, 4, 4, 4), 0)])
def test_steps_vs_oracle(gpu_sq, oracle, dims, pot, real, math):
    rng = np.random.default_rng(4)
    phi0 = rng.normal(size=int(np.prod(dims))) * 0.5
    g, o = pair(gpu_sq, oracle, dims, real, math, pot, m2=0.25, lam=0.5, phi0=phi0)
    for n in (1, 2, 17):
        assert g.step(DTAU, n)
        o.step(DTAU, n)
        m = g.measure()
        assert m["seed"] == o.seed, "step seed must be bit-exact"
        assert m["runs"] == o.L.runs
        err = maxabs(g.download(), o.field)
        assert err < ATOL[(real, math)], err
        tol = 50 * ATOL[(real, math)]
        assert maxabs(m["slice_x"], o.slice_x) < tol and maxabs(m["slice_xx0"], o.slice_xx0) < tol
        fld = o.field.astype(np.float64)
        assert abs(m["mean_phi"] - fld.mean()) < tol and abs(m["mean_phi2"] - (fld ** 2).mean()) < tol
        tm = dims[-1] // 2
        assert maxabs(m["corr"], o.slice_xx0 - o.slice_x * o.slice_x[tm]) < tol
    assert m["nclamped"] == 0


def test_cold_start_c2_shape_small(gpu_sq, oracle):
    """configs[1] at reduced size: cold start phi=0, seed 1242608872, fp32, dtau 0.01."""
    g, o = pair(gpu_sq, oracle, (256, 64), "f32", "fast")
    g.step(DTAU, 40)
    o.step(DTAU, 40)
    assert g.measure()["seed"] == o.seed
    assert maxabs(g.download(), o.field) < ATOL[("f32", "fast")]


@pytest.mark.parametrize("where", ["gid0", "mid", "last", "omega", "plus"])
def test_lattice_rng_events_replayed(gpu_sq, oracle, where):
    """Events are detected on the device, replayed literally on the host, and the step redone:
    seeds bit-exact, fields within tolerance, nevents counted."""
    dims = (32, 16)
    V = 32 * 16
    seed = {"gid0": 177446488061229, "plus": 39512}.get(where)
    if seed is None:
        seed = seed_with_retry_at(oracle, {"mid": 201, "last": V - 1, "omega": V}[where])
    g, o = pair(gpu_sq, oracle, dims, "f64", "accurate", seed=seed)
    g.step(DTAU, 6)
    o.step(DTAU, 6)
    m = g.measure()
    assert m["seed"] == o.seed and o.L.nevents >= 1 and m["nevents"] >= 1
    assert maxabs(g.download(), o.field) < ATOL[("f64", "accurate")]
    assert maxabs(m["slice_x"], o.slice_x) < 1e-5


def test_event_in_later_step(gpu_sq, oracle):
    """An event in step 3 of a 10-step sequence: steps 0-2 stand, 3.. are redone."""
    dims = (16, 8)
    V = 128
    target = seed_with_retry_at(oracle, 77)  # seed at the START of some step
    # walk the chain backwards is hard; instead start 3 steps earlier by brute force over the forward map
    # forward: S_{n+1} = step(S_n).  Choose S_0 freely and check the oracle meets an event later on.
    # invert three whole (event-free) steps of the chain: S_{n+1} = P S_n + c  =>  S_n = (S_{n+1} - c) P^-1
    M = 2**48
    # one full step = V site draws + the omega draw at gid V
    def step_fwd(S):
        s = oracle.lib().sqo_jump(S, 0, V)
        _, rec = oracle.random(s, V)
        return rec.seed_after & (M - 1)
    c = step_fwd(0)
    P = (step_fwd(1) - c) % M
    Pinv = pow(P, -1, M)
    S = target
    for _ in range(3):
        S = ((S - c) * Pinv) % M
    assert step_fwd(step_fwd(step_fwd(S))) == target
    g, o = pair(gpu_sq, oracle, dims, "f32", "accurate", seed=S)
    g.step(DTAU, 10)
    o.step(DTAU, 10)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1
    assert m["seed"] == o.seed and m["runs"] == 10
    assert maxabs(g.download(), o.field) < ATOL[("f32", "accurate")]
    assert maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4


def test_batched_chains(gpu_sq, oracle):
    """configs[4] in miniature: independent chains with their own seeds and couplings."""
    dims, nch = (8, 8, 4, 4), 5
    g = gpu_sq.Context(dims, real="f32", math="accurate", potential=4, nchains=nch, seed=1242608872)
    os_ = []
    rng = np.random.default_rng(8)
    for k in range(nch):
        lam, sd = k / 4.0, 1242608872 + 7 * k
        phi0 = rng.normal(size=int(np.prod(dims))).astype(np.float32) * 0.3
        g.set_chain(k, sd, 0.25, lam)
        g.upload(phi0, chain=k)
        os_.append(oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=lam, seed=sd, phi0=phi0))
    g.step(DTAU, 9)
    mp, mp2, seeds = g.measure_chains()
    for k, o in enumerate(os_):
        o.step(DTAU, 9)
        assert int(seeds[k]) == o.seed
        assert maxabs(g.download(chain=k), o.field) < ATOL[("f32", "accurate")]
        assert abs(mp[k] - o.field.astype(np.float64).mean()) < 1e-5
        assert abs(mp2[k] - (o.field.astype(np.float64) ** 2).mean()) < 1e-5


def test_clamp_counts(gpu_sq):
    """tau_kernel.cl:122-132: an absurd step size blows up; values stay in [-1000, 1000]."""
    g = gpu_sq.Context((32, 32), real="f32", math="fast", potential=0, noise_c=50.0)
    g.step(0.9, 60)
    f = g.download()
    assert np.all(np.isfinite(f)) and np.max(np.abs(f)) <= 1000.0
    assert g.measure()["nclamped"] > 0


def test_full_size_c2_properties(gpu_sq, oracle):
    """configs[1] at full size (1024^2 fp32): step seed bit-exact vs the OpenMP oracle and field parity
    after a few steps; reproducibility (bit-identical reruns); streams independent of launch shape."""
    dims = (1024, 1024)
    g, o = pair(gpu_sq, oracle, dims, "f32", "fast")
    g.step(DTAU, 5)
    o.step(DTAU, 5, omp=True)
    assert g.measure()["seed"] == o.seed
    a = g.download()
    d = np.abs(a.astype(np.float64) - o.field.astype(np.float64))
    assert d.max() < ATOL[("f32", "fast")], d.max()
    assert np.sqrt(np.mean(d ** 2)) < 1e-6  # typical per-site error: fp32 rounding level
    g2 = gpu_sq.Context(dims, real="f32", math="fast")
    g2.step(DTAU, 2); g2.step(DTAU, 3)  # different frame split, same result
    assert np.array_equal(a, g2.download())


def test_free_field_ensemble(gpu_sq):
    """Observables within statistical error: <phi^2> of the 2-D free field vs the analytic Euler-
    discretised value (see tests/test_oracle.py::test_free_field_statistics)."""
    L0 = L1 = 32
    eps = 0.05
    g = gpu_sq.Context((L0, L1), real="f32", math="fast", potential=0)
    g.step(eps, 500)
    vals = []
    for _ in range(400):
        g.step(eps, 10)
        vals.append(g.measure()["mean_phi2"])
    k0 = 2 * np.pi * np.arange(L0) / L0
    lam = 4 * np.sin(k0[:, None] / 2) ** 2 + 4 * np.sin(k0[None, :] / 2) ** 2 + 2.0
    want = float(np.mean(1.0 / (lam * (1 - eps * lam / 2))))
    vals = np.array(vals)
    # binned error (autocorrelation): 20 bins
    bins = vals.reshape(20, -1).mean(axis=1)
    err = bins.std(ddof=1) / np.sqrt(len(bins))
    assert abs(vals.mean() - want) < 5 * err + 1e-3 * want, (vals.mean(), want, err)


# ---------------------------------------------------------------------------------------------
# the on-chip resident kernel (sq_resident.cu): selected automatically for 2-D fp32 lattices with
# rows of 128..1024 sites; same oracle, same tolerances
RES_SHAPES = [(128, 37), (128, 300), (256, 64), (512, 1184), (1024, 592)]


@pytest.mark.parametrize("math", ["fast", "accurate"])
@pytest.mark.parametrize("dims,pot", [(d, p) for d in RES_SHAPES for p in (0, 4)][::2] + [((128, 300), 4), ((1024, 592), 0)])
def test_resident_vs_oracle(gpu_sq, oracle, dims, pot, math):
    rng = np.random.default_rng(12)
    phi0 = (rng.normal(size=int(np.prod(dims))) * 0.5).astype(np.float32)
    g, o = pair(gpu_sq, oracle, dims, "f32", math, pot, m2=0.25, lam=0.5, phi0=phi0)
    for n in (1, 2, 33):
        assert g.step(DTAU, n)
        o.step(DTAU, n, omp=True)
        m = g.measure()
        assert m["seed"] == o.seed and m["runs"] == o.L.runs
        err = maxabs(g.download(), o.field)
        assert err < ATOL[("f32", math)], err
        tol = 50 * ATOL[("f32", math)]
        assert maxabs(m["slice_x"], o.slice_x) < tol and maxabs(m["slice_xx0"], o.slice_xx0) < tol
        tm = dims[-1] // 2
        assert maxabs(m["corr"], o.slice_xx0 - o.slice_x * o.slice_x[tm]) < tol
    assert m["nclamped"] == 0 and m["nevents"] == 0


def test_resident_equals_streaming_stream(gpu_sq, oracle):
    """SQ_FLAG_FORCE_STREAMING: both kernels walk the same integer stream and agree on the field."""
    dims = (256, 96)
    a = gpu_sq.Context(dims, real="f32", math="accurate")
    b = gpu_sq.Context(dims, real="f32", math="accurate", flags=2)
    a.step(DTAU, 21)
    b.step(DTAU, 21)
    assert a.measure()["seed"] == b.measure()["seed"]
    assert maxabs(a.download(), b.download()) < 1e-6
    assert maxabs(a.measure()["slice_xx0"], b.measure()["slice_xx0"]) < 1e-6


@pytest.mark.parametrize("where", ["gid0", "mid", "omega", "plus"])
def test_resident_event_first_step(gpu_sq, oracle, where):
    dims = (128, 40)
    V = 128 * 40
    seed = {"gid0": 177446488061229, "plus": 39512}.get(where)
    if seed is None:
        seed = seed_with_retry_at(oracle, {"mid": 2777, "omega": V}[where])
    g, o = pair(gpu_sq, oracle, dims, "f32", "accurate", seed=seed)
    g.step(DTAU, 12)
    o.step(DTAU, 12)
    m = g.measure()
    assert m["seed"] == o.seed and m["nevents"] >= 1 and o.L.nevents >= 1 and m["runs"] == 12
    assert maxabs(g.download(), o.field) < ATOL[("f32", "accurate")]
    assert maxabs(m["slice_x"], o.slice_x) < 1e-4 and maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4


def test_resident_event_later_step(gpu_sq, oracle):
    """Event in step 5 of a resident launch: the launch is re-run for 5 steps, one streaming step
    takes the event, the resident kernel finishes the frame."""
    dims = (128, 40)
    V = 128 * 40
    M = 2**48
    target = seed_with_retry_at(oracle, 3001)

    def step_fwd(S):
        s = oracle.lib().sqo_jump(S, 0, V)
        _, rec = oracle.random(s, V)
        return rec.seed_after & (M - 1)
    c = step_fwd(0)
    P = (step_fwd(1) - c) % M
    Pinv = pow(P, -1, M)
    S = target
    for _ in range(5):
        S = ((S - c) * Pinv) % M
    g, o = pair(gpu_sq, oracle, dims, "f32", "fast", seed=S)
    g.step(DTAU, 20)
    o.step(DTAU, 20)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1
    assert m["seed"] == o.seed and m["runs"] == 20
    assert maxabs(g.download(), o.field) < ATOL[("f32", "fast")]
    assert maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4


# ---------------------------------------------------------------------------------------------
# BASELINE.json's other configs at full size (SURVEY.md 8(d) C3, C5, and C4's per-GPU slab)
def test_full_size_c3_vs_oracle(gpu_sq, oracle):
    """configs[2] at full size (64^4 fp32, cold start): step seed bit-exact vs the OpenMP oracle, field
    parity after 3 steps, slice observables; the marching kernel is the one that runs."""
    dims = (64, 64, 64, 64)
    g, o = pair(gpu_sq, oracle, dims, "f32", "fast")
    g.step(DTAU, 3)
    o.step(DTAU, 3, omp=True)
    m = g.measure()
    assert m["seed"] == o.seed and m["runs"] == 3
    d = np.abs(g.download().astype(np.float64) - o.field.astype(np.float64))
    assert d.max() < ATOL[("f32", "fast")], d.max()
    assert np.sqrt(np.mean(d ** 2)) < 1e-6
    assert maxabs(m["slice_x"], o.slice_x) < 1e-4 and maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4


def test_c5_chain_grid_properties(gpu_sq, oracle):
    """configs[4] shape (32^4 chains over a lambda grid, 64 of the 512-per-GPU share): a chain's result
    depends only on its own (seed, lambda) -- identical to the same chain run alone -- and two sampled
    chains match the oracle."""
    dims, nch = (32, 32, 32, 32), 64
    lams = np.linspace(0.0, 1.0, 64)
    g = gpu_sq.Context(dims, real="f32", math="fast", potential=4, nchains=nch)
    for k in range(nch):
        g.set_chain(k, 1242608872 + k // 8, 0.25, float(lams[k]))
    g.step(DTAU, 4)
    _, _, seeds = g.measure_chains()
    for k in (0, 37):
        solo = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=float(lams[k]), seed=1242608872 + k // 8)
        solo.step(DTAU, 4)
        assert np.array_equal(g.download(chain=k), solo.download())
        assert int(seeds[k]) == solo.measure()["seed"]
        o = oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=float(lams[k]), seed=1242608872 + k // 8)
        o.step(DTAU, 4, omp=True)
        assert int(seeds[k]) == o.seed
        assert maxabs(g.download(chain=k), o.field) < ATOL[("f32", "fast")]
    # chains with the same seed but different couplings share the noise stream, not the field
    assert int(seeds[0]) == int(seeds[7]) and not np.array_equal(g.download(chain=0), g.download(chain=7))


# ---------------------------------------------------------------------------------------------
# the marching kernel's replay-entry path (virtual start seeds, out-of-line slow strip) with forced events
# at awkward positions: first / last site of a 16-byte strip, row and plane boundaries, last site, omega
MARCH_DIMS = (32, 8, 8, 8)  # L0/4 = 8 threads per row, 32 rows per CTA: takes lattice_march_kernel


@pytest.mark.parametrize("gid", [0, 3, 4, 31, 32, 255, 256, 2047, 2048, 4099, 16383, 16384])
@pytest.mark.parametrize("math", ["fast", "accurate"])
def test_march_kernel_events_replayed(gpu_sq, oracle, gid, math):
    V = int(np.prod(MARCH_DIMS))
    assert gid <= V
    seed = seed_with_retry_at(oracle, gid)
    rng = np.random.default_rng(21)
    phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
    g, o = pair(gpu_sq, oracle, MARCH_DIMS, "f32", math, 4, m2=0.25, lam=0.5, seed=seed, phi0=phi0)
    g.step(DTAU, 5)
    o.step(DTAU, 5)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1
    assert m["seed"] == o.seed and m["runs"] == 5
    assert maxabs(g.download(), o.field) < ATOL[("f32", math)]
    assert maxabs(m["slice_x"], o.slice_x) < 1e-4 and maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4
    assert m["nclamped"] == 0


def test_march_kernel_is_the_one_tested(gpu_sq):
    """Guard for the tests above: same lattice through the generic kernel (SQ_FLAG_GENERIC_KERNEL) walks
    the same stream and agrees on the field to fp32 rounding of the (fused vs separate) noise add."""
    a = gpu_sq.Context(MARCH_DIMS, real="f32", math="accurate", potential=4, m2=0.25, lam=0.5)
    b = gpu_sq.Context(MARCH_DIMS, real="f32", math="accurate", potential=4, m2=0.25, lam=0.5, flags=4)
    a.step(DTAU, 7)
    b.step(DTAU, 7)
    assert a.measure()["seed"] == b.measure()["seed"]
    assert np.array_equal(a.download(), b.download())  # ACCURATE: identical operation sequence in both kernels


def _seed_with_event_in_step(oracle, V, gid, step):
    """Step-start seed S_0 such that the chain meets an inf-retry at `gid` in tau-step `step` (0-based):
    the seed of that step is constructed directly, then `step` whole event-free steps are inverted."""
    M = 2**48
    target = seed_with_retry_at(oracle, gid)

    def step_fwd(S):
        s = oracle.lib().sqo_jump(S, 0, V)
        _, rec = oracle.random(s, V)
        return rec.seed_after & (M - 1)
    c = step_fwd(0)
    P = (step_fwd(1) - c) % M
    Pinv = pow(P, -1, M)
    S = target
    for _ in range(step):
        S = ((S - c) * Pinv) % M
    return S


def test_resident_event_after_a_checkpoint(gpu_sq, oracle):
    """Event in step 300 of a 450-step resident launch: the launch leaves early, the host resumes from the
    checkpoint of step 256 (every 128 steps), reruns 44 steps, takes the event in a streaming step and
    finishes on chip.  Seeds bit-exact, running means over all 450 steps."""
    dims = (128, 40)
    V = 128 * 40
    S = _seed_with_event_in_step(oracle, V, 3001, 300)
    g, o = pair(gpu_sq, oracle, dims, "f32", "fast", seed=S)
    g.step(DTAU, 450)
    o.step(DTAU, 450)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1
    assert m["seed"] == o.seed and m["runs"] == 450
    assert maxabs(g.download(), o.field) < ATOL[("f32", "fast")]
    assert maxabs(m["slice_x"], o.slice_x) < 1e-4 and maxabs(m["slice_xx0"], o.slice_xx0) < 1e-4
    assert m["nclamped"] == 0


def test_march_kernel_3d_event(gpu_sq, oracle):
    dims = (64, 16, 16)  # 16 threads per row, one row per thread: the 3-D instance of the marching kernel
    V = int(np.prod(dims))
    seed = seed_with_retry_at(oracle, 5003)
    g, o = pair(gpu_sq, oracle, dims, "f32", "fast", 0, seed=seed)
    g.step(DTAU, 4)
    o.step(DTAU, 4)
    m = g.measure()
    assert o.L.nevents >= 1 and m["nevents"] >= 1 and m["seed"] == o.seed
    assert maxabs(g.download(), o.field) < ATOL[("f32", "fast")]
    assert m["nclamped"] == 0


def test_event_in_one_chain_of_a_batch(gpu_sq, oracle):
    """Batched chains on the marching kernel: a replay entry belongs to ONE chain; the others must not see it."""
    dims, nch = (32, 8, 8, 8), 3
    seeds = [1242608872, seed_with_retry_at(oracle, 4099), 1242608873]
    g = gpu_sq.Context(dims, real="f32", math="fast", potential=4, nchains=nch)
    for k in range(nch):
        g.set_chain(k, seeds[k], 0.25, 0.5)
    g.step(DTAU, 3)
    _, _, got = g.measure_chains()
    for k in range(nch):
        o = oracle.LatticeOracle(dims, real=oracle.F32, potential=4, m2=0.25, lam=0.5, seed=seeds[k])
        o.step(DTAU, 3)
        assert int(got[k]) == o.seed, k
        assert maxabs(g.download(chain=k), o.field) < ATOL[("f32", "fast")], k
    assert g.measure()["nevents"] >= 1


def test_free_field_ensemble_4d_marching_kernel(gpu_sq):
    """Observables within statistical error on the marching kernel: <phi^2> of the 4-D free field (potID 0:
    F = 2 phi) against the analytic stationary value of the Euler-discretised Langevin process,
    <phi^2> = mean_k 1 / (lam_k (1 - eps lam_k / 2)),  lam_k = sum_mu 4 sin^2(k_mu/2) + 2."""
    L = 16
    eps = 0.05  # stability limit 2 / (16 + 2)
    g = gpu_sq.Context((L, L, L, L), real="f32", math="fast", potential=0)
    g.step(eps, 400)
    vals = []
    for _ in range(240):
        g.step(eps, 5)
        vals.append(g.measure()["mean_phi2"])
    k = 2 * np.pi * np.arange(L) / L
    s2 = 4 * np.sin(k / 2) ** 2
    lam = s2[:, None, None, None] + s2[None, :, None, None] + s2[None, None, :, None] + s2[None, None, None, :] + 2.0
    want = float(np.mean(1.0 / (lam * (1 - eps * lam / 2))))
    vals = np.array(vals)
    bins = vals.reshape(20, -1).mean(axis=1)  # binned error (autocorrelation)
    err = bins.std(ddof=1) / np.sqrt(len(bins))
    assert abs(vals.mean() - want) < 5 * err + 1e-3 * want, (vals.mean(), want, err)


@pytest.mark.parametrize("dims,flag", [((256, 96), 2), ((32, 8, 8, 8), 4), ((64, 16, 16), 4)])
def test_fast_math_is_one_definition_across_kernels(gpu_sq, oracle, dims, flag):
    """SQ_MATH_FAST is the same sequence of operations and roundings in every fp32 lattice kernel (sq_site.cuh:
    site_noise_fast is the scalar form of the packed pipelines), so a run does not depend on which kernel took which
    step: the on-chip kernel / the tile + marching kernels against the generic streaming kernel (SQ_FLAG_FORCE_STREAMING /
    SQ_FLAG_GENERIC_KERNEL), bit for bit -- also across an RNG event, where the replayed step changes kernels."""
    V = int(np.prod(dims))
    rng = np.random.default_rng(31)
    phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
    for seed in (1242608872, seed_with_retry_at(oracle, V // 2 + 3)):
        a = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, seed=seed)
        b = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, seed=seed, flags=flag)
        a.upload(phi0)
        b.upload(phi0)
        a.step(DTAU, 11)
        b.step(DTAU, 11)
        assert a.measure()["seed"] == b.measure()["seed"]
        assert np.array_equal(a.download(), b.download())
        a.close()
        b.close()


@pytest.mark.parametrize("dims,pot", [((32, 32, 8, 6), 0), ((64, 16, 16, 4), 4), ((64, 32, 8), 4), ((256, 16, 8, 4), 0), ((32, 32, 4, 4), 4)])
def test_rowblock_kernel_is_the_tile_kernel_bit_for_bit(gpu_sq, oracle, dims, pot):
    """SQ_FLAG_ROWBLOCK_KERNEL (sq_tile.cu: lattice_rows_kernel -- every stencil operand staged per pass through shared
    memory by a producer warp, tiles claimed from a counter) against the default tile kernel: same integer stream and the
    same field bit for bit, also across an RNG event (the replayed step runs on the marching kernel in both), with several
    chains (tiles of different chains share CTAs) and with the clamp firing.  The observables agree to fp32 summation
    order only: a thread's partial sum runs over a different set of rows in the two kernels."""
    V = int(np.prod(dims))
    rng = np.random.default_rng(32)
    for seed, nch, c in ((1242608872, 3, 1.0), (seed_with_retry_at(oracle, V // 2 + 3), 1, 1.0), (99, 2, 600.0)):
        phi0 = (rng.normal(size=V * nch) * 0.5).astype(np.float32)
        kw = dict(real="f32", math="fast", potential=pot, m2=0.25, lam=0.5, seed=seed, nchains=nch, noise_c=c)
        a = gpu_sq.Context(dims, **kw)
        b = gpu_sq.Context(dims, flags=gpu_sq.SQ_FLAG_ROWBLOCK_KERNEL, **kw)
        for k in range(nch):
            a.upload(phi0[k * V:(k + 1) * V], chain=k)
            b.upload(phi0[k * V:(k + 1) * V], chain=k)
        for _ in range(2):
            a.step(0.9 if c > 1 else DTAU, 7)
            b.step(0.9 if c > 1 else DTAU, 7)
        ma, mb = a.measure(), b.measure()
        assert ma["seed"] == mb["seed"] and ma["nclamped"] == mb["nclamped"]
        if c > 1:
            assert ma["nclamped"] > 0
        for k in ("slice_x", "slice_xx0", "mean_phi", "mean_phi2"):
            x, y = np.asarray(ma[k], dtype=np.float64), np.asarray(mb[k], dtype=np.float64)
            assert np.allclose(x, y, rtol=2e-5, atol=2e-5 * max(1.0, float(np.abs(x).max()))), (k, np.abs(x - y).max())
        for k in range(nch):
            assert np.array_equal(a.download(chain=k), b.download(chain=k))
        a.close()
        b.close()


def test_chain_ensemble_with_jackknife_errors(gpu_sq):
    """stochquant_b200.analysis over batched chains: <phi^2> of the 4-D free field from 24
    independent chains in ONE context, error from a jackknife over chains of bin-averaged series, against the analytic
    stationary value of the Euler-discretised process (see the test above)."""
    from stochquant_b200 import analysis as an
    L, eps, nch = 8, 0.05, 24
    g = gpu_sq.Context((L, L, L, L), real="f32", math="fast", potential=0, nchains=nch)
    g.step(eps, 300)
    _, phi2 = an.collect_series(g, eps, 5, 60)        # [60 frames][24 chains]
    g.close()
    k = 2 * np.pi * np.arange(L) / L
    s2 = 4 * np.sin(k / 2) ** 2
    lam = s2[:, None, None, None] + s2[None, :, None, None] + s2[None, None, :, None] + s2[None, None, None, :] + 2.0
    want = float(np.mean(1.0 / (lam * (1 - eps * lam / 2))))
    per_chain = phi2.mean(axis=0)
    val, err = an.jackknife(lambda a: a.mean(), per_chain)
    assert abs(val - want) < 5 * err + 1e-3 * want, (val, want, err)
    # chains are independent: the scatter over chains matches each chain's own binned error
    _, e1 = an.binned_error(phi2[:, 0], 10)
    assert 0.3 < e1 / per_chain.std(ddof=1) < 3.0


@pytest.mark.parametrize("dims,nsteps", [((256, 96), 40), ((256, 96), 2100), ((32, 8, 8, 8), 6), ((64, 24), 9)])
def test_frame_host_equals_step_and_download(gpu_sq, oracle, dims, nsteps):
    """sq_frame_host (host field in, frame, host field out -- the bench's end-to-end call) reads the field back on its own
    stream right behind the frame's update kernels; the result must be what upload + step + download give, bit for bit:
    without an event, with an RNG event in the frame (the early copy is void and redone), over several frames, and for a
    frame longer than one on-chip launch (2100 steps: no early copy)."""
    V = int(np.prod(dims))
    rng = np.random.default_rng(41)
    phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
    for seed in (1242608872, seed_with_retry_at(oracle, V // 2 + 3)):
        a = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, seed=seed)
        b = gpu_sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, seed=seed)
        hin, hout = phi0.copy(), np.full(V, np.nan, dtype=np.float32)
        a.upload(phi0)
        for _ in range(3):
            a.step(DTAU, nsteps)
            want = a.download()
            ok, _ = b.frame_host(hin.ctypes.data, hout.ctypes.data, DTAU, nsteps)
            assert ok and np.array_equal(hout, want)
            hin[:] = hout
            hout[:] = np.nan
        assert a.measure()["seed"] == b.measure()["seed"]
        a.close()
        b.close()


_GROUPING_SCRIPT = r"""
import hashlib, sys
import numpy as np
import stochquant_b200 as sq
ctx = sq.Context((32, 16, 8, 16), real="f32", math="fast")
h = hashlib.sha256()
for n in (37, 5, 1, 8):   # odd sequence lengths: partial groups, the join at the end of a sequence
    ctx.step(0.01, n)
    m = ctx.measure()
    h.update(ctx.download().tobytes())
    for k in ("slice_x", "slice_xx0", "corr"):
        h.update(np.ascontiguousarray(m[k]).tobytes())
    h.update(repr((m["seed"], m["mean_phi"], m["mean_phi2"], m["runs"], m["nclamped"])).encode())
print("DIGEST", h.hexdigest())
"""


def test_finalize_grouping_and_dependent_launch_change_nothing(gpu_sq):
    """sq_enqueue_step hands the per-step finalize kernels to the side stream in groups of SQ_FIN_BATCH steps (default 4) so
    that the update kernels of a group are neighbours in the stream and overlap by programmatic dependent launch.  That
    only moves work between streams: field, seed and every running mean must be bit-identical to the per-step hand-over
    without dependent launches.  The knobs are read once per process, hence the subprocesses."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for name, env in (("default", {}), ("per_step", {"SQ_FIN_BATCH": "1", "SQ_PDL": "0"}), ("groups_of_8", {"SQ_FIN_BATCH": "8"})):
        e = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), **env)
        for k in ("SQ_FIN_BATCH", "SQ_PDL"):
            if k not in env:
                e.pop(k, None)
        r = subprocess.run([sys.executable, "-c", _GROUPING_SCRIPT], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[name] = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0]
    assert digests["default"] == digests["per_step"] == digests["groups_of_8"], digests
