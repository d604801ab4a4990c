#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/dbg_times.py 10000000 2>&1 | tail -18
timeout -s KILL 300 python tools/dbg_times.py 125000000 2>&1 | tail -9
timeout -s KILL 300 python tools/latency.py 2>&1 | grep " 1000 "
