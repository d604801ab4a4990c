#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 600 python tools/sweep.py 200000000 > gpurun_out/sweep.log 2>&1; grep "rowpop=1" gpurun_out/sweep.log; grep "rowpop=0 warps=16" gpurun_out/sweep.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
export GSB_UNROLL=1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -f -o gpurun_out/prof_scan python tools/prof_driver.py 200000000 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
