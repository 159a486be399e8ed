#!/bin/bash
# Regenerates tests/golden/tauhost_stream_40.txt: the stdout of the drop-in ./tauhost.o on a B200 for the first 40
# frames of taumain.py's own command line (taumain.py:101-132: double_well preset, cold start).  Run on the GPU box:
#     gpurun -- 'bash tests/golden/make_tauhost_stream.sh && cp /tmp/tauhost_stream_40.txt gpurun_out/'
set -e
cd "$(dirname "$0")/../.."
./tauhost.o 200 0.02 0.002 40 3 1.0 2 1 0 1000 0 /tmp/V0_2e_0-8.txt 40 > /tmp/tauhost_stream_40.txt
wc -l /tmp/tauhost_stream_40.txt
