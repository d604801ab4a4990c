#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/dbg_times.py 10000000 2>&1 | tail -20
timeout -s KILL 300 python tools/dbg_times.py 125000000 2>&1 | tail -20
GSB_ONLY=1 timeout -s KILL 600 python tools/sweep.py 200000000 2>&1 | grep "rowpop=1 warps=1[26] stages=2 U=1"
