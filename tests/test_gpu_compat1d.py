"""GPU parity: the 1-D reference-faithful kernel (SQ_KERNEL_COMPAT1D) through the C-ABI against
the oracle (canonical {chain, Jacobi} restatement of /root/reference/tau_kernel.cl:25-175).

Tiers (BASELINE.json north_star): integer RNG stream, seeds, site indexing, lrgEl, step counts
and commit/rollback are BIT-EXACT; fp64 field values agree to ATOL below -- the only
difference is CUDA's cosf/logf/tanhf vs glibc's (<= 2 ulp of fp32 each, i.e. ~1e-7 on the
N(0,1) draw, times the noise amplitude ~0.1, accumulated over the steps of a test)."""
import ctypes as C

import numpy as np
import pytest

from helpers import maxabs, seed_with_retry_at

pytestmark = pytest.mark.gpu

ATOL = 2e-6  # absolute, fp64 fields of O(1) after <= 10^3 steps


def make(gpu_sq, oracle, N=200, dt=.02, dtau=1e-4, pot=3, c=1.0, seed=None):
    f, om, r1 = oracle.host_init(N, dt, dtau)
    if seed is not None:
        r1 = seed
    g = gpu_sq.Context([N], kernel="compat1d", potential=pot, spacing=dt, noise_c=c, f0=f, omega0=om, seed=r1)
    o = oracle.Compat1D(N, dt, dtau, pot, c, f, om, r1)
    return g, o


def test_integer_stream_bit_exact(gpu_sq, oracle):
    """(t1,t2) of every work-item of the first step == tau_kernel.cl:273,275 in gid order."""
    g, o = make(gpu_sq, oracle)
    t1, t2 = g.debug_draws(0, 201)
    trace = (oracle.Draw * 201)()
    oracle.lib().sqo_set_trace(trace, 201)
    o.launch(1)
    oracle.lib().sqo_set_trace(None, 0)
    assert [int(v) for v in t1] == [d.t1 for d in trace]
    assert [int(v) for v in t2] == [d.t2 for d in trace]


@pytest.mark.parametrize("pot,N,dt,dtau", [(3, 200, .02, 1e-4), (0, 100, .1, 3e-3), (3, 17, .05, 5e-4),
                                           (3, 1000, .02, 1e-4), (0, 3000, .05, 1e-3)])
def test_frame_vs_oracle(gpu_sq, oracle, pot, N, dt, dtau):
    g, o = make(gpu_sq, oracle, N, dt, dtau, pot)
    for loops in (1, 7, 60):
        assert g.step(dtau, loops) is True
        assert o.frame(loops) is True
        m = g.measure()
        assert m["seed"] == o.s.rand1, "RNG seed after the frame must be bit-exact"
        assert m["lrgEl"] == o.s.lrgEl
        assert m["steps_done"] == loops and m["runs"] == o.s.runs
        assert maxabs(m["f"], o.f) < ATOL
        assert maxabs(m["x"], o.x) < ATOL and maxabs(m["xx0"], o.xx0) < ATOL
        assert abs(m["omega"] - o.s.omega) < ATOL and abs(m["lrgVl"] - o.s.lrgVl) < ATOL
        xavg = o.xx0 - o.x * o.x[N // 2]
        assert maxabs(m["corr"], xavg) < 2 * ATOL
        cl = np.array([oracle.lib().sqo_clas(i * dt, o.s.omega, pot) for i in range(N)])
        assert abs(m["mean_phi"] - np.mean(o.f + cl)) < ATOL
        assert abs(m["mean_phi2"] - np.mean((o.f + cl) ** 2)) < ATOL


def test_unstable_frame_rolls_back(gpu_sq, oracle):
    """tauhost.c:533-554: state restored, but seed / lrgEl / lrgVl keep advancing."""
    dtau = 2e-3  # 10x above the Euler limit of the default run
    g, o = make(gpu_sq, oracle, dtau=dtau)
    f0 = o.f.copy()
    n_exec = o.launch(50)
    assert o.s.stable == 0 and n_exec < 50
    assert g.step(dtau, 50) is False
    m = g.measure()
    assert m["steps_done"] == n_exec
    assert np.array_equal(m["f"], f0) and not m["x"].any()
    assert m["seed"] == o.s.rand1 and m["lrgEl"] == o.s.lrgEl
    assert abs(m["lrgVl"] - o.s.lrgVl) < 1e-6 * max(1.0, o.s.lrgVl)
    assert m["runs"] == 0


def test_controller_sequence(gpu_sq, oracle):
    """The host's dtau controller (tauhost.c:506-545) driven by the GPU's stable flags follows
    the oracle's accept/reject sequence exactly."""
    N, dt, dtau0, loops = 200, .02, .002, 40
    g, o = make(gpu_sq, oracle, dtau=dtau0)
    dg, do_, cg, co = dtau0, dtau0, 0, 0
    seq_g, seq_o = [], []
    for _ in range(70):
        o.s.deltaTau = do_
        so = o.frame(loops)
        sg = g.step(dg, loops)
        seq_g.append(sg); seq_o.append(so)
        if so:
            if co > 10: co, do_ = 0, do_ / 0.95
            co += 1
        else:
            do_, co = do_ * 0.95, 0
        if sg:
            if cg > 10: cg, dg = 0, dg / 0.95
            cg += 1
        else:
            dg, cg = dg * 0.95, 0
    assert seq_g == seq_o and dg == do_
    assert 5 < seq_g.count(False) < 70
    m = g.measure()
    assert m["seed"] == o.s.rand1 and m["runs"] == o.s.runs
    assert maxabs(m["f"], o.f) < 1e-4  # near the stability edge errors are amplified


def test_rng_events_replayed(gpu_sq, oracle):
    """inf-retry (:282) and `seed+=` (:278-279) inside a frame: seeds stay bit-exact."""
    cases = [177446488061229,            # retry at gid 0 of the first step
             39512,                      # seed < 2^31 and t2 < 2^31 at gid 0: `*seed += temp`
             seed_with_retry_at(oracle, 137),   # retry in the middle of the lattice
             seed_with_retry_at(oracle, 200),   # retry on the omega work-item
             2**64 - 5]                  # u64 wrap of seed+gid
    for seed in cases:
        g, o = make(gpu_sq, oracle, seed=seed)
        assert g.step(1e-4, 25) and o.frame(25)
        m = g.measure()
        assert m["seed"] == o.s.rand1, hex(seed)
        assert maxabs(m["f"], o.f) < ATOL
        if seed != 2**64 - 5:
            assert m["nevents"] >= 1


def test_runs_offset_and_restart_state(gpu_sq, oracle):
    """runs0 (the `runs` kernel argument) enters the Welford denominators (:144-145)."""
    N, dt, dtau = 64, .05, 4e-4
    f, om, r1 = oracle.host_init(N, dt, dtau)
    rng = np.random.default_rng(2)
    x0, xx00 = rng.normal(size=N) * .1, rng.normal(size=N) * .01
    g = gpu_sq.Context([N], kernel="compat1d", potential=3, spacing=dt, f0=f, x0=x0, xx0_0=xx00, omega0=om, seed=r1)
    o = oracle.Compat1D(N, dt, dtau, 3, 1.0, f, om, r1, x=x0, xx0=xx00, runs=5000)
    assert g.step(dtau, 30, runs0=5000) and o.frame(30)
    m = g.measure()
    assert maxabs(m["x"], o.x) < ATOL and maxabs(m["xx0"], o.xx0) < ATOL and m["runs"] == 5030


# ---------------------------------------------------------------------------------------------
# f-3: the frame controller on the device (sq_controller_* / sq_frames) against the same rules applied
# by the host frame by frame (host/tauhost.c, tauhost.c:504-545): BIT-identical trajectories
@pytest.mark.parametrize("pot,N,dt,dtau0", [(3, 200, .02, .002), (0, 100, .1, 3e-3)])
def test_device_controller_equals_host_loop(gpu_sq, oracle, pot, N, dt, dtau0):
    """The preset step size is above the Euler limit on purpose (taumain.py:102-104): the first frames
    are rejected and dtau shrinks, later it grows again -- every branch of the controller runs."""
    f, om, r1 = oracle.host_init(N, dt, dtau0)
    loops, frames = 50, 90
    a = gpu_sq.Context([N], kernel="compat1d", potential=pot, spacing=dt, f0=f, omega0=om, seed=r1)
    b = gpu_sq.Context([N], kernel="compat1d", potential=pot, spacing=dt, f0=f, omega0=om, seed=r1)
    # host loop
    dtau, runs, stab = dtau0, 0, 0
    host = []
    for _ in range(frames):
        st = a.step(dtau, loops, runs)
        host.append((dtau, int(st), a.measure()["steps_done"]))
        if st:
            if stab > 10:
                stab, dtau = 0, dtau / 0.950
            stab += 1
            runs += loops
        else:
            dtau, stab = dtau * 0.950, 0
    # device controller, in batches that do not divide the frame count
    b.controller_set(dtau0, 0, 0)
    dev, xav = [], []
    left = frames
    while left:
        n = min(left, 37)
        r, x = b.frames(n, loops)
        dev += r
        xav += list(x)
        left -= n
    assert [h[1] for h in host] == [d[1] for d in dev], "accept / reject sequence"
    assert [h[0] for h in host] == [d[0] for d in dev], "dtau sequence must be bit-identical"
    assert [h[2] for h in host] == [d[2] for d in dev]
    assert any(d[1] == 1 for d in dev) and (pot != 3 or any(d[1] == 0 for d in dev))
    ma, mb = a.measure(), b.measure()
    for k in ("f", "x", "xx0"):
        assert np.array_equal(ma[k], mb[k]), k
    assert ma["seed"] == mb["seed"] and ma["omega"] == mb["omega"] and ma["lrgEl"] == mb["lrgEl"]
    assert b.controller_get() == (dtau, runs, stab)
    # the logged xavg of the last accepted frame is the host's xx0 - x*x[mid]
    last = max(k for k, d in enumerate(dev) if d[1] == 1)
    if last == frames - 1:
        assert np.array_equal(xav[last], ma["xx0"] - ma["x"] * ma["x"][N // 2])


def test_frames_argument_checks(gpu_sq, oracle):
    f, om, r1 = oracle.host_init(50, .05, 1e-3)
    g = gpu_sq.Context([50], kernel="compat1d", potential=0, spacing=.05, f0=f, omega0=om, seed=r1)
    g.controller_set(1e-3, 0, 0)
    recs, _ = g.frames(0, 10)
    assert recs == []
    with pytest.raises(gpu_sq.SqError):
        g.frames(gpu_sq.SQ_FRAMES_MAX + 1, 10)
    with pytest.raises(gpu_sq.SqError):
        g.controller_set(0.0, 0, 0)
    recs, xav = g.frames(gpu_sq.SQ_FRAMES_MAX, 5)  # a full log
    assert len(recs) == gpu_sq.SQ_FRAMES_MAX and g.controller_get()[1] == 5 * sum(r[1] for r in recs)
