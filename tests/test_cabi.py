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
