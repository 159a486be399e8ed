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
