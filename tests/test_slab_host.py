"""CPU (no GPU): the host side of the multi-GPU slab ring, world_size 2.

  * the session (shared-memory barrier + all-gathers, stochquant_b200/csrc/sq_session.cu) against
    torch.distributed's gloo collectives run side by side;
  * the per-step event agreement of sq_slab.cu::resolve_step, mirrored here rank by rank with the
    oracle's literal chain standing in for the CUDA finder kernel: every rank must end the step with
    the reference's seed (tau_kernel.cl:269-284 in gid order) wherever the event sits.
"""
import ctypes as C
import os
import sys
import uuid

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M48 = 2**48 - 1
NONE = 2**64 - 1


def _scan(O, sq, base_seed, base_gid, lo, hi):
    """Stand-in for find_events_kernel: first gid in [lo, hi) whose literal draw is an event."""
    if lo >= hi:
        return NONE
    s = C.c_uint64(base_seed if lo == base_gid else sq.load().sq_lcg_jump(base_seed, base_gid, lo - base_gid))
    for g in range(lo, hi):
        rec = O.Draw()
        O.lib().sqo_random(C.byref(s), g, C.byref(rec))
        if rec.ndraws > 1 or rec.plus_branch:
            return g
    return NONE


def _rank_main(rank, world, port, name, seeds, V, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import stochquant_b200 as sq
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sess = sq.Session(name, rank, world)
        # ---- session collectives vs gloo --------------------------------------------------------
        for it in range(50):
            mine = np.array([rank * 1000 + it, it * it + rank], dtype=np.uint64)
            got = sess.allgather_u64(mine)
            ref = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(ref, torch.from_numpy(mine.astype(np.int64)))
            assert np.array_equal(got.astype(np.int64), torch.stack(ref).numpy())
        x = np.arange(300, dtype=np.float64) * (rank + 1)
        got = sess.allgather_f64(x)
        ref = [torch.zeros(300, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(ref, torch.from_numpy(x))
        assert np.array_equal(got, torch.stack(ref).numpy())
        sess.barrier()
        # ---- event agreement: rank r scans gids [lo, hi) of every step -------------------------
        lo, hi = (0, V // 2) if rank == 0 else (V // 2, V)
        finals = []
        for S in seeds:
            entries, bg, bs, frm = [], 0, S, lo
            while True:
                local = _scan(O, sq, bs, bg, frm, hi)
                allv = sess.allgather_u64([local])[:, 0]
                t = torch.tensor([local if local != NONE else -1], dtype=torch.int64)
                ref = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
                dist.all_gather(ref, t)
                assert [int(v) if int(v) != NONE else -1 for v in allv] == [int(r.item()) for r in ref]
                g = int(allv.min())
                if g == NONE:
                    break
                e, nd, plus = sq.rng_resolve(S, entries, g)
                assert nd > 1 or plus
                entries.append(e)
                bg, bs = g + 1, e["seed"]
                frm = max(g + 1, lo)
            e, nd, plus = sq.rng_resolve(S, entries, V)  # the omega work-item's draw
            finals.append((e["seed"], len(entries) + (1 if (nd > 1 or plus) else 0)))
        sess.close()
        q.put((rank, finals, None))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc() + repr(ex)))
    finally:
        dist.destroy_process_group()


def test_session_and_event_agreement_world2(oracle, sq):
    import torch.multiprocessing as mp
    from helpers import seed_with_retry_at
    V = 2048
    # events in rank 0's range, in rank 1's range, at the last site, at the omega draw, the
    # `seed+=` branch at gid 0, and an event-free step
    seeds = [seed_with_retry_at(oracle, 300), seed_with_retry_at(oracle, 1500), seed_with_retry_at(oracle, V - 1),
             seed_with_retry_at(oracle, V), 39512, 1242608872]
    want = []
    for S in seeds:
        o = oracle.LatticeOracle((64, V // 64), real=oracle.F64, seed=S)
        o.step(0.01, 1)
        want.append((o.seed, int(o.L.nevents)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    name = "h" + uuid.uuid4().hex[:12]
    ps = [ctx.Process(target=_rank_main, args=(r, 2, port, name, seeds, V, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, finals, err in out:
        assert err is None, err
        for (seed, nev), (wseed, wnev) in zip(finals, want):
            assert seed == wseed, (rank, hex(seed), hex(wseed))
            assert nev == wnev
    assert want[-1][1] == 0 and all(w[1] >= 1 for w in want[:-1])


def test_session_rejects_bad_arguments(sq):
    L = sq.load()
    h = C.c_void_p()
    assert L.sq_session_open(C.byref(h), b"x/y", 0, 1) != 0
    assert L.sq_session_open(C.byref(h), b"ok", 2, 2) != 0
    assert L.sq_session_open(C.byref(h), b"ok", 0, 1000) != 0
    s = sq.Session("solo" + uuid.uuid4().hex[:8], 0, 1)  # a ring of one needs nobody else
    assert np.array_equal(s.allgather_u64([7, 8]), np.array([[7, 8]], dtype=np.uint64))
    s.barrier()
    s.close()


def test_session_missing_rank_times_out(sq, tmp_path):
    """A ring whose partner never shows up must fail (SQ_ERR_TIMEOUT), not hang."""
    import subprocess, sys, time
    code = ("import sys; sys.path.insert(0, %r); import stochquant_b200 as sq\n"
            "try:\n    sq.Session('lonely' + %r, 0, 2)\n    print('opened')\n"
            "except sq.SqError as e:\n    print('error', e.code)\n") % (ROOT, uuid.uuid4().hex[:8])
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60,
                       env=dict(os.environ, SQ_SESSION_TIMEOUT="1.5"))
    assert r.stdout.strip() == "error -6", (r.stdout, r.stderr)
    assert time.time() - t0 < 30
