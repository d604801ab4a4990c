#!/bin/bash
# tensor-core kernel: parity tests, timings at 100 M and 1 B rows, role clocks
set -u
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_tensor.py -x -q -m gpu --timeout 100 > gpurun_out/pytest_tensor.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_tensor.log
timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 3 2>&1 | tail -3 | head -2
GSB_TC_DEBUG=1 timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
timeout -s KILL 120 python tools/prof_tensor.py 1000000000 128 2 2>&1 | tail -3
