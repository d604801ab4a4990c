#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 600 python tools/sweep.py 200000000 > gpurun_out/sweep.log 2>&1; grep "rowpop=1" gpurun_out/sweep.log; grep "rowpop=0 warps=16\|rowpop=0 warps=12 stages=2" gpurun_out/sweep.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
