"""CPU: the C-ABI library loads, exports every symbol include/sq.h declares, and refuses to
run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(sq):
    L = sq.load()
    names = sq.exported_symbols()
    assert {"sq_init", "sq_step", "sq_measure", "sq_free"} <= set(names)
    nm = subprocess.run(["nm", "-D", "--defined-only", sq.library_path()], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (sq_[a-z0-9_]+)", nm))
    assert set(names) <= defined, sorted(set(names) - defined)
    assert L.sq_api_version() == 3


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sq.h"\nint main(void){sq_params p; sq_obs o; (void)p; (void)o; return sizeof(p)>0?0:1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layout_matches_ctypes(sq, tmp_path):
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "sq.h"\nint main(void){printf("%zu %zu\\n", sizeof(sq_params), sizeof(sq_obs));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b = map(int, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert a == C.sizeof(sq.SqParams) and b == C.sizeof(sq.SqObs)


def test_no_cpu_fallback(sq):
    """Without a CUDA device every entry into the compute path must fail loudly."""
    L = sq.load()
    if L.sq_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(sq.SqError) as e:
        sq.Context([8, 8])
    assert e.value.code == -5
    with pytest.raises(sq.SqError):
        sq.Context([200], kernel="compat1d", potential=3, spacing=.02)


def test_bad_arguments_rejected(sq):
    L = sq.load()
    assert L.sq_init(None, None, None, None, None, 0.0, 0) == -1
    p = sq.SqParams()
    p.struct_size = 3  # wrong size
    h = C.c_void_p()
    assert L.sq_init(C.byref(h), C.byref(p), None, None, None, 0.0, 0) == -1
    assert b"no CPU fallback" in L.sq_strerror(-5)
    assert L.sq_step(None, 0.1, 1, 0, None) == -1
    assert L.sq_measure(None, None) == -1
    L.sq_free(None)


def test_product_never_references_oracle():
    """The shipped code must not include, link or import anything under oracle/."""
    bad = []
    for base in ("stochquant_b200", "host", "include"):
        for dp_, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".cu", ".cuh", ".h", ".c", ".py")):
                    txt = open(os.path.join(dp_, fn)).read()
                    if re.search(r"sq_oracle|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(dp_, fn))
    assert not bad, bad
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "liboracle" not in mk and "lsq_oracle" not in mk


def test_lcg_jump_host_utility(sq, oracle):
    """sq_lcg_jump (product, host arithmetic) == oracle's independent jump-ahead == literal chain."""
    import random
    L = sq.load()
    rnd = random.Random(9)
    for _ in range(5000):
        s, g = rnd.getrandbits(48), rnd.getrandbits(rnd.choice([4, 12, 20, 33]))
        D = rnd.getrandbits(rnd.choice([1, 8, 17, 33, 39]))
        assert L.sq_lcg_jump(s, g, D) == oracle.lib().sqo_jump(s, g, D)
    s = C.c_uint64(1242608872)
    for g in range(2000):
        if g % 211 == 0:
            assert L.sq_lcg_jump(1242608872, 0, g) == (s.value & (2**48 - 1))
        oracle.lib().sqo_random(C.byref(s), g, None)


def test_rng_resolve_against_the_survey_kats(sq, oracle):
    """sq_rng_resolve is the PRODUCT's literal replay of one draw (the host side of every RNG-event recovery and of the
    ring's agreement, tau_kernel.cl:269-284 incl. the do/while retry and the `seed +=` branch).  Pinned here, without a
    GPU, to SURVEY.md 8(c) G1 -- the same integers the oracle is pinned to -- and to the oracle on random inputs."""
    def at(seed, gid):  # the draw at `gid` from `seed`: an entry that says "draws at gid >= gid continue from seed"
        e = {"gid_start": gid, "seed": seed, "ov_gid": 2**64 - 1, "ov_t1": 0, "ov_t2": 0}
        return sq.rng_resolve(0, [e] if gid else [], gid) if gid else sq.rng_resolve(seed, [], 0)
    plain = [(1242608872, 0, 0x8f7a818576d3, 0x9dba931929e2, 173422509894114),
             (1242608872, 1, 0x8f8060725d40, 0x58e126661ab8, 97721887627960),
             (1242608872, 200, 0x9410aa997bfb, 0xd3dda7355112, 232946799038738),
             (1, 0, 0x5deece678, 0xbb61488df123, 206024356000035),
             (2**64 - 5, 7, 0xbbdd9cce5, 0x76ab15684887, 130475023157383)]
    for s, g, t1, t2, nx in plain:
        e, nd, plus = at(s, g)
        assert (e["ov_gid"], e["ov_t1"], e["ov_t2"], e["seed"], e["gid_start"], nd, plus) == (g, t1, t2, nx, g + 1, 1, 0)
    for s, g, t1, t2, nx in [(1760221443, 3, 277355843144601, 121207, 1760342650), (668289095, 3, 64439088464781, 140379, 668429474)]:
        e, nd, plus = at(s, g)   # `*seed += temp` (:278-279)
        assert (e["ov_t1"], e["ov_t2"], e["seed"], nd, plus) == (t1, t2, nx, 1, 1)
    for s, g, t1, t2, nx in [(177446488061229, 0, 0xdc5d786b660e, 0xf1c4fc530801, 0xf1c47c530801),
                             (177446488061224, 5, 0x841e58ec1a3c, 0xc0818e0993b8, 0xc0810e0993b8)]:
        e, nd, plus = at(s, g)   # inf-retry (:282): the first t1 is 0x1234, the draw is repeated from the updated seed
        assert (e["ov_t1"], e["ov_t2"], e["seed"], nd) == (t1, t2, nx, 2)
    import random
    rnd = random.Random(5)
    for _ in range(2000):
        s, g = rnd.getrandbits(rnd.choice([20, 31, 48, 64])), rnd.getrandbits(rnd.choice([1, 8, 20, 33]))
        e, nd, plus = at(s, g)
        _, rec = oracle.random(s, g)
        assert (e["ov_t1"], e["ov_t2"], e["seed"], nd, plus) == (rec.t1, rec.t2, rec.seed_after, rec.ndraws, rec.plus_branch)
