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
