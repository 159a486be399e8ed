"""GPU: the drop-in binary ./tauhost.o end to end -- argv of taumain.py:132, stdout stream parsed the
way taumain.py:27-41 parses it, end file -- against the whole reference program restated on the
CPU (oracle.tauhost_main = tauhost.c:29-621 over the oracle kernel, canonical semantics)."""
import os
import re
import subprocess
from io import BytesIO

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tauhost.o")


def parse_stream(raw: bytes):
    """taumain.py:30-41: per line genfromtxt(delimiter='|'); [-2]=dtau, [-1]=percent, rest = plot data."""
    frames = []
    for line in raw.splitlines():
        tmp = np.genfromtxt(BytesIO(line.strip()), delimiter="|")
        assert tmp.size > 1
        frames.append((tmp[:-2], tmp[-2], tmp[-1]))
    return frames


def run_pair(oracle, tmp_path, args, cwd):
    out_o = tmp_path / "oracle_stdout.txt"
    end_o, end_g = tmp_path / "end_oracle.txt", tmp_path / "end_gpu.txt"
    a_o = list(args); a_o[11] = str(end_o)
    a_g = list(args); a_g[11] = str(end_g)
    assert oracle.tauhost_main(a_o, str(out_o)) == 0
    r = subprocess.run([EXE] + [str(a) for a in a_g], capture_output=True, cwd=cwd, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stderr == b""
    return r.stdout, out_o.read_bytes(), end_g.read_text(), end_o.read_text()


def test_default_preset_short(gpu_sq, oracle, tmp_path):
    """taumain.py's double_well preset (n=200, dt=.02, dtau=.002, potID 3, loops 1000) for 60 frames;
    runs in an empty cwd: no tau_kernel.cl needed (tauhost.c:187-191 is gone)."""
    args = [200, 0.02, 0.002, 60, 3, 1.0, 2, 1, 0, 1000, "0", "end.txt", 40]
    so, ref, end_g, end_o = run_pair(oracle, tmp_path, args, cwd=str(tmp_path))
    fg, fo = parse_stream(so), parse_stream(ref)
    assert len(fg) == len(fo) == 60
    assert so.splitlines()[0] == ref.splitlines()[0]  # first line: all -inf, dtau, percent -- byte-identical
    for (yg, dg, pg), (yo, do_, po) in zip(fg, fo):
        assert yg.size == 199 and dg == do_ and pg == po  # same accept/reject + dtau controller path
    # the data column is log|xavg|: compare xavg itself, the controller parks the run near the Euler limit
    xg, xo = np.exp(fg[-1][0]), np.exp(fo[-1][0])
    assert np.max(np.abs(xg - xo)) < 1e-5
    # end file: identical layout; numbers close; trailer identical where it is integer / controller state
    lg, lo = end_g.split("\n"), end_o.split("\n")
    assert len(lg) == len(lo) == 204
    assert lg[201] == lo[201] and lg[202] == lo[202]  # N line and deltaTau line
    pat = re.compile(r"^[ -]0x[01]\.?[0-9a-f]*p[+-]\d+ *\|omega$")
    assert pat.match(lg[200])
    vg = np.array([[float.fromhex(t.strip()) for t in l.split("|")] for l in lg[:200]])
    vo = np.array([[float.fromhex(t.strip()) for t in l.split("|")] for l in lo[:200]])
    assert all(len(l) == 4 * 40 + 3 * 2 for l in lg[:200])  # 4 fields of width 40 joined by "| "
    assert np.max(np.abs(vg - vo)) < 1e-4


def test_harmosc_preset_and_restart(gpu_sq, oracle, tmp_path):
    """harmosc preset (taumain.py:92-100) then a restart from its end file (tauhost.c:103-173)."""
    args = [100, 0.1, 0.3, 25, 0, 1.0, 0, 1, 0, 200, "0", "e.txt", 30]
    so, ref, end_g, end_o = run_pair(oracle, tmp_path, args, cwd=str(tmp_path))
    fg, fo = parse_stream(so), parse_stream(ref)
    assert [f[1] for f in fg] == [f[1] for f in fo]
    # restart both from the ORACLE's end file so the inputs are identical
    start = tmp_path / "start.txt"
    start.write_text(end_o)
    args2 = [100, 0.1, 0.3, 10, 0, 1.0, 0, 2, 0, 200, str(start), "e2.txt", 30]
    so2, ref2, end_g2, end_o2 = run_pair(oracle, tmp_path, args2, cwd=str(tmp_path))
    fg2, fo2 = parse_stream(so2), parse_stream(ref2)
    assert len(fg2) == len(fo2) == 5  # fps=2: every second frame
    assert [f[1] for f in fg2] == [f[1] for f in fo2]
    assert end_g2.split("\n")[101] == end_o2.split("\n")[101]  # the (double-counting) N line, tauhost.c:577
    x0 = np.exp(fg2[0][0]); x1 = np.exp(fo2[0][0])
    assert np.allclose(x0, x1, rtol=0, atol=1e-9)  # first line prints the restart file's xavg


def test_windows_15_arg_form(gpu_sq, tmp_path):
    """taumain_windows.py:163 passes (n, dt, dtau, h, parisi, frames, potID, ...)."""
    a13 = ["50", "0.1", "0.001", "4", "3", "1.0", "0", "1", "0", "20", "0", "0", "40"]
    a15 = a13[:3] + ["1.0", "0"] + a13[3:]
    r13 = subprocess.run([EXE] + a13, capture_output=True, cwd=str(tmp_path), timeout=120)
    r15 = subprocess.run([EXE] + a15, capture_output=True, cwd=str(tmp_path), timeout=120)
    assert r13.returncode == 0 and r15.returncode == 0
    assert r13.stdout == r15.stdout and r13.stdout.count(b"\n") == 4


def test_unwritable_end_file(gpu_sq, tmp_path):
    a = ["50", "0.1", "0.001", "2", "3", "1.0", "0", "1", "0", "5", "0", "/nonexistent_dir/out.txt", "40"]
    r = subprocess.run([EXE] + a, capture_output=True, cwd=str(tmp_path), timeout=120)
    assert r.returncode == 1 and r.stderr == b"Failed to write to Output.\n"  # tauhost.c:565-566


def test_extended_trailer_resumes_bit_exactly(gpu_sq, tmp_path):
    """SURVEY.md 8(f) f-1.  With TAUHOST_EXT_TRAILER=1 a run stopped after K frames and restarted from its end file
    for K more leaves the SAME end file, byte for byte, as one run of 2K frames -- seed, omega, lrgEl / lrgVl, the
    stale newf[lrgEl], the step size to the last bit and the controller's counter all travel in the trailer.  (The
    double-well preset rejects ~45 frames first, so rollbacks, step-size changes and the stale newf[lrgEl] are all
    exercised.)  Without the variable the restart behaves as the reference does (re-randomised omega and seed):
    the end files then differ."""
    K = 70
    env = dict(os.environ, TAUHOST_EXT_TRAILER="1")
    base = ["200", "0.02", "0.002", None, "3", "1.0", "2", "1", "0", "1000", None, None, "40"]

    def run(frames, start, end, e):
        a = list(base)
        a[3], a[10], a[11] = str(frames), start, str(end)
        r = subprocess.run([EXE] + a, capture_output=True, cwd=str(tmp_path), timeout=600, env=e)
        assert r.returncode == 0, r.stderr.decode()
        return r.stdout

    whole, half, rest = tmp_path / "whole.txt", tmp_path / "half.txt", tmp_path / "rest.txt"
    out_whole = run(2 * K, "0", whole, env)
    run(K, "0", half, env)
    out_rest = run(K, str(half), rest, env)
    assert whole.read_bytes() == rest.read_bytes()
    assert b"|sqext" in whole.read_bytes()
    # the frame lines of the second half show the same data and step sizes (the percent column restarts)
    lw, lr = out_whole.splitlines()[K:], out_rest.splitlines()
    assert len(lw) == len(lr) == K
    for u, v in zip(lw, lr):
        assert u.rsplit(b"|", 1)[0] == v.rsplit(b"|", 1)[0]
    # reference behaviour without the variable: the same restart re-randomises -> a different end file
    plain = tmp_path / "plain.txt"
    run(K, str(half), plain, dict(os.environ))
    assert plain.read_bytes() != rest.read_bytes() and b"|sqext" not in plain.read_bytes()


def test_stream_matches_golden(gpu_sq, tmp_path):
    """The committed frame stream that tests/test_taumain_headless.py feeds to the reference's unmodified
    front-end IS what the drop-in binary writes for taumain.py's own command line (first 40 frames): same
    accept / reject sequence and step sizes, same data to 1e-6 (CUDA's logf / cosf / tanhf may move by an ulp
    between toolkits, and log|xavg| of a value near zero amplifies it)."""
    golden = open(os.path.join(ROOT, "tests", "golden", "tauhost_stream_40.txt"), "rb").read()
    args = ["200", "0.02", "0.002", "40", "3", "1.0", "2", "1", "0", "1000", "0", str(tmp_path / "end.txt"), "40"]
    r = subprocess.run([EXE] + args, capture_output=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    fg, fo = parse_stream(r.stdout), parse_stream(golden)
    assert len(fg) == len(fo) == 40
    assert r.stdout.splitlines()[0] == golden.splitlines()[0]
    for (yg, dg, pg), (yo, do_, po) in zip(fg, fo):
        assert dg == do_ and pg == po and yg.size == yo.size == 199
        fin = np.isfinite(yo)
        assert np.array_equal(fin, np.isfinite(yg))
        assert np.allclose(np.exp(yg[fin]), np.exp(yo[fin]), rtol=0, atol=1e-6)
