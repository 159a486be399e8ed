"""CPU: the drop-in host's non-GPU pieces (host/tauhost_io.c) against the oracle's restatement of
/root/reference/tauhost.c:84-102 (initial state), :485-501 (stdout line), :562-581 / :103-173
(end / start file), byte for byte."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def io():
    so = os.path.join(ROOT, "host", "libtauhost_io.so")
    subprocess.run(["make", "-C", ROOT, "host/libtauhost_io.so"], check=True, capture_output=True)
    L = C.CDLL(so)
    pd = C.POINTER(C.c_double)
    L.th_initial_state.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, pd, pd, C.POINTER(C.c_ulong)]
    L.th_write_end_file.argtypes = [C.c_char_p, C.c_int, C.c_int, pd, pd, pd, pd, C.c_double, C.c_int, C.c_double]
    L.th_read_start_file.argtypes = [C.c_char_p, C.c_int, C.c_double, pd, pd, pd, pd, C.POINTER(C.c_int), pd]
    L.th_write_end_file_ext.argtypes = [C.c_char_p, C.c_int, C.c_int, pd, pd, pd, pd, C.c_double, C.c_int, C.c_double, C.POINTER(ThExt)]
    L.th_read_start_file_ext.argtypes = [C.c_char_p, C.c_int, C.c_double, pd, pd, pd, pd, C.POINTER(C.c_int), pd,
                                         C.POINTER(ThExt), C.POINTER(C.c_int)]
    return L


class ThExt(C.Structure):  # host/tauhost_io.h: th_ext
    _fields_ = [("seed", C.c_ulonglong), ("lrgEl", C.c_int), ("stab_cnt", C.c_int), ("runs", C.c_longlong),
                ("lrgVl", C.c_double), ("omega", C.c_double), ("newf_lrgEl", C.c_double), ("dtau", C.c_double)]


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


HELPER = r'''
#include <stdio.h>
#include <stdlib.h>
#include "tauhost_io.h"
int main(int argc, char **argv) {
    int n = atoi(argv[1]); double dt = atof(argv[2]), dtau = atof(argv[3]);
    double *f = calloc(n, sizeof(double)), om; unsigned long r1;
    th_initial_state(n, dt, dtau, 1, f, &om, &r1);
    printf("%a %lu", om, r1);
    for (int i = 0; i < n; ++i) printf(" %a", f[i]);
    printf("\n");
    return 0;
}
'''


def test_initial_state_fresh_process(oracle, tmp_path):
    """rand() state is per process: run th_initial_state in a fresh process like tauhost.o does."""
    src = tmp_path / "h.c"
    src.write_text(HELPER)
    exe = tmp_path / "h"
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "host"), str(src),
                    os.path.join(ROOT, "host", "tauhost_io.c"), "-lm", "-o", str(exe)], check=True)
    for n, dt, dtau in [(200, .02, .002), (100, .1, .3), (7, .5, .01)]:
        out = subprocess.run([str(exe), str(n), str(dt), str(dtau)], check=True, capture_output=True, text=True).stdout.split()
        f, om, r1 = oracle.host_init(n, dt, dtau)
        assert float.fromhex(out[0]) == om and int(out[1]) == r1
        assert np.array_equal(np.array([float.fromhex(t) for t in out[2:]]), f)


def test_stdout_line_bytes(io, oracle, tmp_path):
    rng = np.random.default_rng(5)
    for n in (2, 5, 200):
        xavg = rng.normal(size=n) * 10.0 ** rng.integers(-12, 3, size=n)
        xavg[1 % n] = 0.0  # log(0) = -inf
        if n > 3:
            xavg[3] = np.nan
        a, b = tmp_path / "a.txt", tmp_path / "b.txt"
        fp = C.CDLL(None).fopen(os.fsencode(str(a)), b"w")
        # use the product function through a FILE*: simplest is a tiny C shim via libc fopen
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fclose.argtypes = [C.c_void_p]
        fp = libc.fopen(os.fsencode(str(a)), b"w")
        io.th_print_frame.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int]
        io.th_print_frame(fp, n, dp(xavg), 1.9888e-4, 41, 5000)
        libc.fclose(fp)
        oracle.lib().sqo_print_frame_path(os.fsencode(str(b)), n, dp(xavg), 1.9888e-4, 41, 5000)
        assert a.read_bytes() == b.read_bytes()
        if n == 200:
            line = a.read_text()
            assert line.count("|") == n and line.endswith(" 0.84\n")


def test_appendix_a_formats(io, tmp_path):
    """literal examples of the end-file lines."""
    n = 1
    xavg, xx0, x, f = (np.array([v]) for v in (0.46430999999999998, -0.25, 0.0, 0.001))
    p = tmp_path / "end.txt"
    assert io.th_write_end_file(os.fsencode(str(p)), n, 40, dp(xavg), dp(xx0), dp(x), dp(f),
                                float.fromhex("0x1.02f28b90dc1b3p+1"), 248000, 1.9888e-4) == 0
    lines = p.read_text().split("\n")
    assert lines[0] == " 0x1.db7414a4d2b2cp-2                   | -0x1p-2                                 |  0x0p+0                                 |  0x1.0624dd2f1a9fcp-10                  "
    assert lines[1] == " 0x1.02f28b90dc1b3p+1                   |omega"
    assert lines[2] == "                                  248000|N"
    assert lines[3] == " 1.988800e-04                           |deltaTau"


def test_end_file_bytes_and_round_trip(io, oracle, tmp_path):
    rng = np.random.default_rng(11)
    n = 50
    xavg, xx0, x, f = (rng.normal(size=n) for _ in range(4))
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    for acc in (40, 25):
        assert io.th_write_end_file(os.fsencode(str(a)), n, acc, dp(xavg), dp(xx0), dp(x), dp(f), 2.0231, 123000, 1.5e-4) == 0
        assert oracle.lib().sqo_write_endfile(os.fsencode(str(b)), n, acc, dp(xavg), dp(xx0), dp(x), dp(f), 2.0231, 123000, 1.5e-4) == 0
        assert a.read_bytes() == b.read_bytes()
    # round trip: %a is exact; omega line ignored; N line -> rec; dtau capped at the CLI value
    for reader, path in ((io.th_read_start_file, a), (oracle.lib().sqo_read_startfile, b)):
        r = [np.zeros(n) for _ in range(4)]
        rec, dt = C.c_int(-1), C.c_double(-1.0)
        assert reader(os.fsencode(str(path)), n, 1e-3, dp(r[0]), dp(r[1]), dp(r[2]), dp(r[3]), C.byref(rec), C.byref(dt)) == 0
        assert all(np.array_equal(u, v) for u, v in zip(r, (xavg, xx0, x, f)))
        assert rec.value == 123000 and dt.value == 1.5e-4
        dt2 = C.c_double(-1.0)
        reader(os.fsencode(str(path)), n, 1e-4, dp(r[0]), dp(r[1]), dp(r[2]), dp(r[3]), C.byref(rec), C.byref(dt2))
        assert dt2.value == 1e-4  # tauhost.c:133-135
    assert io.th_read_start_file(b"/nonexistent/file", n, 1e-3, dp(xavg), dp(xx0), dp(x), dp(f), C.byref(rec), C.byref(dt)) == 1
    assert io.th_write_end_file(b"/nonexistent/dir/file", n, 40, dp(xavg), dp(xx0), dp(x), dp(f), 0., 0, 0.) == 1


def test_extended_trailer_round_trip_and_reference_compat(io, oracle, tmp_path):
    """the extended trailer sits BEHIND the reference's three trailer lines.  It round-trips
    every field exactly (hex floats, u64 seed), the plain reader and the oracle's restatement of the reference's
    reader (tauhost.c:116-168) read such a file exactly as if the trailer were not there, and a file without (or
    with a truncated) trailer reports has_ext = 0."""
    rng = np.random.default_rng(3)
    n = 37
    xavg, xx0, x, f = (rng.normal(size=n) for _ in range(4))
    ext = ThExt(seed=2**64 - 12345, lrgEl=17, stab_cnt=7, runs=123456789012, lrgVl=1.0 / 3.0, omega=np.nextafter(2.0231, 3),
                newf_lrgEl=-0.1 / 7.0, dtau=1.5e-4 / 0.95 / 0.95)
    a, b = tmp_path / "ext.txt", tmp_path / "plain.txt"
    p = lambda q: os.fsencode(str(q))
    assert io.th_write_end_file_ext(p(a), n, 40, dp(xavg), dp(xx0), dp(x), dp(f), 2.0231, 123000, 1.5e-4, C.byref(ext)) == 0
    assert io.th_write_end_file_ext(p(b), n, 40, dp(xavg), dp(xx0), dp(x), dp(f), 2.0231, 123000, 1.5e-4, None) == 0
    ta, tb = a.read_bytes(), b.read_bytes()
    assert ta.startswith(tb) and ta[len(tb):].startswith(b"1|sqext\n") and ta.count(b"\n") == tb.count(b"\n") + 9
    oracle.lib().sqo_write_endfile(p(tmp_path / "o.txt"), n, 40, dp(xavg), dp(xx0), dp(x), dp(f), 2.0231, 123000, 1.5e-4)
    assert tb == (tmp_path / "o.txt").read_bytes()  # ext == NULL is the reference's file, byte for byte
    got, has = ThExt(), C.c_int(-1)
    r = [np.zeros(n) for _ in range(4)]
    rec, dt = C.c_int(-1), C.c_double(-1.0)
    assert io.th_read_start_file_ext(p(a), n, 1e-3, dp(r[0]), dp(r[1]), dp(r[2]), dp(r[3]), C.byref(rec), C.byref(dt), C.byref(got), C.byref(has)) == 0
    assert has.value == 1
    for k, _ in ThExt._fields_:
        assert getattr(got, k) == getattr(ext, k), k
    assert all(np.array_equal(u, v) for u, v in zip(r, (xavg, xx0, x, f))) and rec.value == 123000 and dt.value == 1.5e-4
    # the reference's reader (oracle restatement) and the plain reader ignore the trailer
    for reader in (io.th_read_start_file, oracle.lib().sqo_read_startfile):
        r2 = [np.zeros(n) for _ in range(4)]
        rec2, dt2 = C.c_int(-1), C.c_double(-1.0)
        assert reader(p(a), n, 1e-3, dp(r2[0]), dp(r2[1]), dp(r2[2]), dp(r2[3]), C.byref(rec2), C.byref(dt2)) == 0
        assert all(np.array_equal(u, v) for u, v in zip(r2, (xavg, xx0, x, f))) and rec2.value == 123000 and dt2.value == 1.5e-4
    # no trailer / truncated trailer -> has_ext = 0
    for path, blob in ((b, tb), (tmp_path / "cut.txt", ta[:ta.rfind(b"\n", 0, len(ta) - 1) + 1])):
        path.write_bytes(blob)
        has.value = -1
        assert io.th_read_start_file_ext(p(path), n, 1e-3, dp(r[0]), dp(r[1]), dp(r[2]), dp(r[3]), C.byref(rec), C.byref(dt), C.byref(got), C.byref(has)) == 0
        assert has.value == 0


def test_cli_forms_and_errors(tmp_path):
    """13-arg (tauhost.c:31-43) and 15-arg (taumain_windows.py:163) forms; bad potIDs; no-GPU exit."""
    exe = os.path.join(ROOT, "tauhost.o")
    subprocess.run(["make", "-C", ROOT, "tauhost.o"], check=True, capture_output=True)
    r = subprocess.run([exe, "200"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    base = ["200", "0.02", "0.002", "2", "1", "1.0", "2", "1", "0", "10", "0", "0", "40"]
    r = subprocess.run([exe] + base, capture_output=True, text=True)
    assert r.returncode == 1 and "potID 1 is not supported" in r.stderr and r.stdout == ""
    win = ["200", "0.02", "0.002", "1.0", "1", "2", "3", "1.0", "0", "1", "0", "10", "0", "0", "40"]
    r = subprocess.run([exe] + win, capture_output=True, text=True)
    assert r.returncode == 2 and "parisi" in r.stderr
    base[4] = "3"
    base[10] = str(tmp_path / "missing_start.txt")
    r = subprocess.run([exe] + base, capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr == "Failed to read Input.\n"  # tauhost.c:106-107
