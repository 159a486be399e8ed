"""CPU: the reference's UNMODIFIED front-end (/root/reference/taumain.py, run where it lies) against the frame
stream of the drop-in ./tauhost.o.

taumain.py spawns `./tauhost.o` from the working directory with its 13 positional arguments (:132), parses every
stdout line with np.genfromtxt(delimiter='|') (:27-41) and animates it through matplotlib (:62-89).  There is no
matplotlib and no GPU in this container, and no /root/reference on the GPU box, so the two halves meet here:
  * `./tauhost.o` is a replayer of tests/golden/tauhost_stream_40.txt -- the byte stream the REAL drop-in binary
    wrote on a B200 for the first 40 frames of taumain.py's own command line (tests/golden/make_tauhost_stream.sh;
    tests/test_gpu_tauhost.py::test_stream_matches_golden regenerates it on the GPU box and compares);
  * matplotlib is the headless stub in tests/stubs/ whose show() drives the animation callbacks.
The script must run to completion, hand every frame's 199 values to the plot, and report the last step size."""
import json
import os
import stat
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAUMAIN = "/root/reference/taumain.py"
GOLDEN = os.path.join(ROOT, "tests", "golden", "tauhost_stream_40.txt")


@pytest.mark.skipif(not os.path.exists(TAUMAIN), reason="the reference front-end is only present in the build container")
def test_unmodified_taumain_runs_headless_on_the_dropin_stream(tmp_path):
    # a stand-in ./tauhost.o that checks the argv taumain.py:132 passes and replays the recorded stream
    fake = tmp_path / "tauhost.o"
    fake.write_text(f"""#!{sys.executable}
import sys, time
a = sys.argv[1:]
assert a == ['200', '0.02', '0.002', '5000', '3', '1.0', '2', '1', '0', '1000', '0', 'V0_2e_0-8.txt', '40'], a
for line in open({GOLDEN!r}, 'rb'):
    sys.stdout.buffer.write(line)
    sys.stdout.buffer.flush()
    time.sleep(0.002)
open(a[11], 'w').write('end file placeholder\\n')
""")
    fake.chmod(fake.stat().st_mode | stat.S_IXUSR)
    report = tmp_path / "report.json"
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "stubs"), SQ_MPL_STUB_REPORT=str(report),
               SQ_MPL_STUB_TIMEOUT="120")
    r = subprocess.run([sys.executable, TAUMAIN], cwd=str(tmp_path), env=env, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    out = r.stdout.decode()
    # taumain.py:43-45: the data thread's last status line; the step size is the stream's last dtau field
    stream = open(GOLDEN, "rb").read().splitlines()
    assert len(stream) == 40
    last = np.genfromtxt([stream[-1].strip().decode()], delimiter="|")
    assert "100.00%" + "| DeltaTau = %.2e| " % last[-2] in out
    rep = json.loads(report.read_text())
    assert rep["npoints"] == 199 and rep["ylim"] == [-15.0, 15.0]  # taumain.py:137, :128
    assert rep["distinct_frames"] >= 2 and rep["updates"] >= rep["distinct_frames"]
    # the plot ended on a frame of the stream (the animation thread takes what the queue holds when it polls)
    ys = [np.genfromtxt([l.strip().decode()], delimiter="|")[:-2] for l in stream]
    got = np.array(rep["last_y"])
    assert any(np.array_equal(np.nan_to_num(got, nan=-1e300, neginf=-1e308), np.nan_to_num(y, nan=-1e300, neginf=-1e308)) for y in ys)
    assert (tmp_path / "V0_2e_0-8.txt").exists()


TAUMAIN_WIN = "/root/reference/taumain_windows.py"
# what taumain_windows.py:163 passes for its selected preset (`double_well`, :14, :129-160):
# n, deltat, deltatau, h, parisi, entw, potID, c, device, rpf, intime, loops, inputf, outputf, acco
WIN_ARGV = ['100', '1.0', '0.01', '1e-05', '0', '100', '3', '1.0', '0', '1', '0', '10000', '0', '0', '40']


@pytest.mark.skipif(not os.path.exists(TAUMAIN_WIN), reason="the reference front-end is only present in the build container")
def test_unmodified_taumain_windows_runs_headless_with_its_15_arguments(tmp_path):
    """The other front-end north_star names: /root/reference/taumain_windows.py, run where it lies, unmodified.  It spawns
    `tauhost.exe` (looked up on PATH) with FIFTEEN positional arguments (:163), reads stdout byte by byte (:33) and plots
    from a thread (:60-97).  The stand-in `tauhost.exe` checks that argv, and -- as there is no GPU here -- writes the
    frame stream of the same run with 20 instead of 10000 tau-steps per frame through the CPU restatement of the
    reference host (oracle.tauhost_main; same line format as the drop-in, tests/test_host_io.py).  That the REAL drop-in
    accepts exactly this argv is asserted on its parser: without a GPU it gets as far as sq_init and exits 3, not 2."""
    stream = tmp_path / "stream.txt"
    fake = tmp_path / "tauhost.exe"
    fake.write_text(f"""#!{sys.executable}
import sys, time
sys.path.insert(0, {ROOT!r})
a = sys.argv[1:]
assert a == {WIN_ARGV!r}, a
from oracle import oracle as O
n, dt, dtau, h, parisi, frames, pot, c, dev, fps, intime, loops, fin, fout, acc = a
assert O.tauhost_main([n, dt, dtau, frames, pot, c, dev, fps, intime, '20', fin, fout, acc], {str(stream)!r}) == 0
for line in open({str(stream)!r}, 'rb'):
    sys.stdout.buffer.write(line)
    sys.stdout.buffer.flush()
    time.sleep(0.002)
""")
    fake.chmod(fake.stat().st_mode | stat.S_IXUSR)
    report = tmp_path / "report.json"
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "stubs"), SQ_MPL_STUB_REPORT=str(report),
               SQ_MPL_STUB_TIMEOUT="120", PATH=str(tmp_path) + os.pathsep + os.environ.get("PATH", ""))
    r = subprocess.run([sys.executable, TAUMAIN_WIN], cwd=str(tmp_path), env=env, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert "100.00%, " in r.stdout.decode()                       # taumain_windows.py:56: the data thread saw EOF and left
    lines = open(stream, "rb").read().splitlines()
    assert len(lines) == 100                                     # entw frames, rpf = 1
    rep = json.loads(report.read_text())
    assert rep["npoints"] == 99 and rep["ylim"] == [-15.0, 15.0]  # :174 (range(n-1)), :160 (theoVal 10 -> dmax 15)
    assert rep["distinct_frames"] >= 2
    ys = [np.genfromtxt([l.strip().decode()], delimiter="|")[:-2] for l in lines]
    got = np.array(rep["last_y"])
    assert any(np.array_equal(np.nan_to_num(got, nan=-1e300, neginf=-1e308), np.nan_to_num(y, nan=-1e300, neginf=-1e308)) for y in ys)
    # the drop-in's own parser takes this argv (15-argument form: h ignored, parisi = 0) and only then looks for a GPU
    exe = os.path.join(ROOT, "tauhost.o")
    if os.path.exists(exe):
        import stochquant_b200 as sq
        if sq.load().sq_device_count() <= 0:
            rr = subprocess.run([exe] + WIN_ARGV, capture_output=True, text=True, timeout=120)
            assert rr.returncode == 3 and rr.stdout == "" and "usage" not in rr.stderr, (rr.returncode, rr.stderr)
