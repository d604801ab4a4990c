#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 2>&1 | tail -3 | head -2
GSB_TC_DEBUG=1 timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
for f in 6; do
GSB_TC_FAULT=$f timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 2>&1 | tail -3 | head -2
GSB_TC_DEBUG=1 GSB_TC_FAULT=$f timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
done
