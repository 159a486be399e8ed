"""CPU: the oracle restatement against the reference's own outputs.

Golden vectors in tests/golden/ were produced by the reference kernel source itself
(tests/golden/make_golden.py over oracle/_ref); where oracle/_ref is present (dev container)
the comparison is also made live.  Cites /root/reference/tau_kernel.cl.
"""
import ctypes as C
import os

import numpy as np
import pytest


def test_random_kats_bit_exact(oracle, golden_dir):
    """tau_kernel.cl:269-284: value, and seed after, for 2500 (seed, gid) pairs."""
    g = np.load(os.path.join(golden_dir, "random_kat.npz"))
    for s, gid, v, after in zip(g["seed"], g["gid"], g["value"], g["seed_after"]):
        val, rec = oracle.random(int(s), int(gid))
        assert rec.seed_after == int(after)
        assert val == v or (np.isnan(val) and np.isnan(v))


def test_survey_integer_kats(oracle):
    """SURVEY.md 8(c) G1, re-derived: (seed,gid) -> (t1,t2,next seed)."""
    kats = [(1242608872, 0, 0x8f7a818576d3, 0x9dba931929e2, 173422509894114),
            (1242608872, 1, 0x8f8060725d40, 0x58e126661ab8, 97721887627960),
            (1242608872, 200, 0x9410aa997bfb, 0xd3dda7355112, 232946799038738),
            (1, 0, 0x5deece678, 0xbb61488df123, 206024356000035),
            (2**64 - 5, 7, 0xbbdd9cce5, 0x76ab15684887, 130475023157383)]
    for s, g, t1, t2, nx in kats:
        _, rec = oracle.random(s, g)
        assert (rec.t1, rec.t2, rec.seed_after, rec.ndraws, rec.plus_branch) == (t1, t2, nx, 1, 0)
    # `*seed += temp` branch (:278-279)
    for s, g, t1, t2, nx in [(1760221443, 3, 277355843144601, 121207, 1760342650),
                             (668289095, 3, 64439088464781, 140379, 668429474)]:
        _, rec = oracle.random(s, g)
        assert (rec.t1, rec.t2, rec.seed_after, rec.plus_branch) == (t1, t2, nx, 1)
    # inf-retry (:282): first t1 = 0x1234 -> v1 = 0 -> redraw with the updated seed
    _, rec = oracle.random(177446488061229, 0)
    assert (rec.t1, rec.t2, rec.seed_after, rec.ndraws) == (0xdc5d786b660e, 0xf1c4fc530801, 0xf1c47c530801, 2)
    _, rec = oracle.random(177446488061224, 5)
    assert (rec.t1, rec.t2, rec.seed_after, rec.ndraws) == (0x841e58ec1a3c, 0xc0818e0993b8, 0xc0810e0993b8, 2)
    _, rec = oracle.random(0, 0)  # t1 = 0xb -> retry
    assert rec.ndraws == 2


def test_chain_default_run_step0(oracle):
    """seeds after gid 0,1,2,200 of step 0 of the default run."""
    s = C.c_uint64(1242608872)
    want = {0: 173422509894114, 1: 244753361787202, 2: 208390381140088, 200: 260177099296138}
    for g in range(201):
        oracle.lib().sqo_random(C.byref(s), g, None)
        if g in want:
            assert s.value == want[g]


def test_glibc_initial_state(oracle):
    """tauhost.c:84-102,185 with the unseeded glibc rand() stream."""
    f, om, r1 = oracle.host_init(200, .02, .002)
    assert om.hex() == "0x1.02f28b90dc1b3p+1"
    assert f[0].hex() == "0x1.aec813f42a4eep-7" and f[1].hex() == "0x1.20ff40b281d6dp-7"
    assert f[199].hex() == "-0x1.c463b4d8ebfa7p-4"
    assert r1 == 1242608872
    _, om2, r2 = oracle.host_init(100, .1, .3)
    assert om2.hex() == "0x1.520ca06ee145ap+2" and r2 == 1939964443


def test_model_functions(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "model_fns.npz"))
    L = oracle.lib()
    cl = np.array([L.sqo_clas(a, w, 3) for a, w in zip(g["a"], g["w"])])
    assert np.array_equal(cl, g["clas3"])
    assert np.array_equal(np.array([L.sqo_ddPot(x, 3) for x in cl]), g["ddpot3"])
    assert L.sqo_intConst(0) == g["intconst"][0] and L.sqo_intConst(3) == g["intconst"][1]
    assert L.sqo_clas(1.3, 0.2, 0) == 0.0 and L.sqo_ddPot(0.7, 0) == 2.0


CASES = ["dw200", "ho100", "dw17", "dw200_unstable"]


@pytest.mark.parametrize("name", CASES)
def test_time_dev_canonical_vs_reference_golden(oracle, golden_dir, name):
    """{chain, Jacobi}: the oracle equals the reference kernel run as Loops=1 launches, bit for bit."""
    g = np.load(os.path.join(golden_dir, "time_dev_ref.npz"))
    pot, N, dt, dtau, steps = g[name + "_params"]
    pot, N, steps = int(pot), int(N), int(steps)
    o = oracle.Compat1D(N, dt, dtau, pot, 1.0, g[name + "_f0"], float(g[name + "_omega0"][0]),
                        int(g[name + "_seed0"][0]))
    seeds = []
    for _ in range(len(g[name + "_seeds"])):
        o.launch(1)
        seeds.append(o.s.rand1)
        if o.s.stable != 1:
            break
        o.f[:], o.x[:], o.xx0[:] = o.newf, o.newx, o.newxx0
        o.s.runs += 1
    assert np.array_equal(np.array(seeds, dtype=np.uint64), g[name + "_seeds"])
    assert np.array_equal(o.newf, g[name + "_newf"])
    assert np.array_equal(o.newx, g[name + "_newx"])
    assert np.array_equal(o.newxx0, g[name + "_newxx0"])
    assert o.s.omega == g[name + "_omega"][0]
    assert [o.s.lrgEl, o.s.lrgVl, o.s.stable] == list(g[name + "_lrg"])


@pytest.mark.parametrize("name", [c for c in CASES if "unstable" not in c])
def test_time_dev_inplace_vs_reference_golden(oracle, golden_dir, name):
    """{chain, in-place}: one launch with Loops=steps under the serial schedule.

    Only for stable runs: after a step sets *stable=0 the reference has a FIFTH race -- the
    `if(*stable!=1) break` after the barrier (tau_kernel.cl:168-171) is read by work-items
    while earlier ones already execute the next step; under a serial work-item schedule items
    behind the one that tripped the flag skip that step (207 instead of 402 draws in the
    dw200_unstable fixture).  Canonical semantics = lock-step: every item finishes the step
    in which the flag fell, then all break (test_time_dev_canonical_* covers that)."""
    g = np.load(os.path.join(golden_dir, "time_dev_ref.npz"))
    pot, N, dt, dtau, steps = g[name + "_params"]
    o = oracle.Compat1D(int(N), dt, dtau, int(pot), 1.0, g[name + "_f0"], float(g[name + "_omega0"][0]),
                        int(g[name + "_seed0"][0]), field=oracle.FIELD_INPLACE)
    o.launch(int(steps))
    assert np.array_equal(o.newf, g[name + "_inplace_newf"])
    assert np.array_equal(o.newx, g[name + "_inplace_newx"])
    assert o.s.rand1 == int(g[name + "_inplace_seed"][0])
    assert [o.s.lrgEl, o.s.lrgVl, o.s.stable] == list(g[name + "_inplace_lrg"])


def test_live_reference_when_available(oracle):
    """Dev container only: compare against oracle/_ref built from /root/reference right now."""
    if not os.path.isdir("/root/reference") or not oracle.ref_available():
        pytest.skip("reference tree not present (GPU box): golden fixtures cover this")
    rng = np.random.default_rng(3)
    R = oracle.ref()
    for _ in range(3000):
        s, g = int(rng.integers(0, 2**48)), int(rng.integers(0, 8192))
        a = C.c_ulong(s)
        v = R.sq_ref_random(C.byref(a), g)
        v2, rec = oracle.random(s, g)
        assert v == v2 and a.value == rec.seed_after
    f, om, r1 = oracle.host_init(64, .05, 4e-4)
    o = oracle.Compat1D(64, .05, 4e-4, 3, 0.7, f, om, r1)
    r = oracle.RefKernel(64, .05, 4e-4, 3, 0.7, f, om, r1)
    for _ in range(30):
        o.frame(1)
        r.steps_canonical(1)
    assert np.array_equal(o.f, r.f) and np.array_equal(o.xx0, r.xx0) and o.s.rand1 == r.rand1.value


def test_jump_ahead_matches_literal_chain(oracle):
    """sqo_jump (used by the OpenMP baseline) == literal sequential chain where no event occurs."""
    s = C.c_uint64(1242608872)
    s0 = s.value
    M = 2**48 - 1
    for g in range(3000):
        if g % 97 == 0:
            assert oracle.lib().sqo_jump(s0, 0, g) == (s.value & M)
        rec = oracle.Draw()
        oracle.lib().sqo_random(C.byref(s), g, C.byref(rec))
        assert rec.ndraws == 1 and rec.plus_branch == 0
    # from a non-zero base gid
    base = s.value
    for g in range(3000, 3500):
        oracle.lib().sqo_random(C.byref(s), g, None)
    assert oracle.lib().sqo_jump(base, 3000, 500) == (s.value & M)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("dims", [(8, 6), (8, 4, 6), (4, 4, 4, 6)])
def test_lattice_omp_equals_serial(oracle, real, dims):
    """The OpenMP/jump-ahead CPU baseline equals the serial definition bit for bit."""
    r = oracle.F32 if real == "f32" else oracle.F64
    rng = np.random.default_rng(1)
    phi0 = rng.normal(size=int(np.prod(dims)))
    a = oracle.LatticeOracle(dims, real=r, potential=4, m2=0.25, lam=0.5, phi0=phi0)
    b = oracle.LatticeOracle(dims, real=r, potential=4, m2=0.25, lam=0.5, phi0=phi0)
    a.step(0.01, 5)
    b.step(0.01, 5, omp=True)
    assert np.array_equal(a.field, b.field) and a.seed == b.seed
    assert np.allclose(a.slice_x, b.slice_x, rtol=0, atol=1e-14)
    assert np.allclose(a.slice_xx0, b.slice_xx0, rtol=0, atol=1e-14)


def test_omp_thread_count_is_settable_and_result_independent(oracle):
    """bench.py's CPU legs set the thread count themselves (torchrun exports OMP_NUM_THREADS=1);
    the field does not depend on it."""
    prev = oracle.set_threads(0)
    outs = []
    for n in (1, 2, 5):
        assert oracle.set_threads(n) == n
        o = oracle.LatticeOracle((16, 6, 10), real=oracle.F32, potential=4, m2=0.25, lam=0.5)
        o.step(0.01, 4, omp=True)
        outs.append((o.field.copy(), o.seed))
    oracle.set_threads(prev)
    assert all(np.array_equal(outs[0][0], f) and outs[0][1] == sd for f, sd in outs[1:])


def test_lattice_event_replay_serial_vs_omp(oracle):
    """A seed that hits the inf-retry inside the lattice: both paths must agree (OMP falls back)."""
    # seed 177446488061229 retries at gid 0
    for seed in (177446488061229, 1760221443 - 0):
        a = oracle.LatticeOracle((8, 8), seed=seed)
        b = oracle.LatticeOracle((8, 8), seed=seed)
        a.step(0.01, 3)
        b.step(0.01, 3, omp=True)
        assert np.array_equal(a.field, b.field) and a.seed == b.seed
    assert a.L.nevents >= 0


def test_free_field_statistics(oracle):
    """Analytic check of the 2-D free field: <phi^2> of the Euler-discretised
    Langevin process, phi' = (1 - eps K) phi + sqrt(2 eps) r, with K's eigenvalues
    lam_k = sum_mu 4 sin^2(k_mu/2) + 2 and stationary variance sum_k 1/(V lam_k (1 - eps lam_k / 2))."""
    L0, L1, eps = 16, 16, 0.05
    o = oracle.LatticeOracle((L0, L1), real=oracle.F64, potential=0)
    o.step(eps, 600)  # thermalise
    acc, n = 0.0, 0
    for _ in range(3000):
        o.step(eps, 1)
        acc += float(np.mean(o.field.astype(np.float64) ** 2))
        n += 1
    k0 = 2 * np.pi * np.arange(L0) / L0
    k1 = 2 * np.pi * np.arange(L1) / L1
    lam = 4 * np.sin(k0[:, None] / 2) ** 2 + 4 * np.sin(k1[None, :] / 2) ** 2 + 2.0
    want = float(np.mean(1.0 / (lam * (1 - eps * lam / 2))))
    assert abs(acc / n - want) < 0.03 * want


@pytest.mark.parametrize("a", [1.0, 0.5, 0.25, 2.0])
@pytest.mark.parametrize("real", ["f64", "f32"])
def test_lattice_definition_reduces_to_the_reference_kernel_in_1d(oracle, a, real):
    """The d-dimensional update has no reference code; its DEFINITION (sqo_lattice_step) is pinned
    here to the reference kernel: a 1 x N lattice (one site per time slice, the x0 neighbours are the site itself) is the
    reference's chain away from its ends -- same shared-seed draw per site (gid = t, the omega item's draw at gid = N), same
    Box-Muller, same Euler step (tau_kernel.cl:114 with potID 0: W = 2, clas = 0).  Compared with the restatement of
    time_dev (bit-equal to the reference's own source by the tests above) and, where oracle/_ref exists, with that
    source itself.  The two differ only (i) at the ends (reference: fixed ghosts, lattice: periodic), which move inwards
    one site per step -- the comparison keeps clear of them -- and (ii) in the noise amplitude, sqrt(2 dtau / a^d) with
    d = 2 here against d = 1, undone through C = sqrt(a) up to the float rounding of the amplitude.  Observables: the
    slice means of a 1 x N lattice are the reference's x_i and xx0_i (:144-145)."""
    N, dtau, K, seed = 200, 0.002, 20, 1242608872
    f0 = 0.1 * np.random.default_rng(7).standard_normal(N)
    if real == "f32":
        f0 = f0.astype(np.float32).astype(np.float64)
    chains = [oracle.Compat1D(N, a, dtau, 0, 1.0, f0, 0.0, seed)]
    if oracle.ref_available():
        chains.append(oracle.RefKernel(N, a, dtau, 0, 1.0, f0, 0.0, seed))
    lat = oracle.LatticeOracle((1, N), real=oracle.F64 if real == "f64" else oracle.F32, potential=0, a=a,
                               c=float(np.sqrt(a)), seed=seed, phi0=f0)
    for _ in range(K):
        lat.step(dtau)
        for c in chains:
            if isinstance(c, oracle.Compat1D):
                assert c.frame(1)
                assert c.s.rand1 == lat.seed            # the integer stream: bit-exact, step by step
            else:
                assert c.steps_canonical(1)
                assert c.rand1.value == lat.seed
    lo, hi = K + 2, N - K - 2
    # fp64: one rounding of the amplitude in float (2^-24 relative of a noise term ~0.1 per step); fp32 storage: 2^-24 of phi per step
    tol = 2e-7 if real == "f64" else 2e-5
    for c in chains:
        assert np.abs(lat.field[lo:hi].astype(np.float64) - c.f[lo:hi]).max() < tol
        assert np.abs(lat.slice_x[lo:hi] - c.x[lo:hi]).max() < tol
        assert np.abs(lat.slice_xx0[lo:hi] - c.xx0[lo:hi]).max() < tol
    if a in (1.0, 0.25) and real == "f64":  # amplitudes that are exact in float: equal to the last bit or two
        assert np.abs(lat.field[lo:hi] - chains[0].f[lo:hi]).max() < 1e-15
