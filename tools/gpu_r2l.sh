#!/bin/bash
set -u
mkdir -p gpurun_out
GSB_TC_DEBUG=1 timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
GSB_TC_DEBUG=1 timeout -s KILL 600 python tools/prof_tensor.py 400000000 128 2 2>&1 | tail -8 | head -7
