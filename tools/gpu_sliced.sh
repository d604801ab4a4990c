#!/bin/bash
# Bit-sliced multi-query kernel: parity tests, batch bench (all variants), ncu capture.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -k "sliced or multi_query" --timeout 600 -x > gpurun_out/pytest_sliced.log 2>&1
echo "pytest sliced rc=$?"; tail -5 gpurun_out/pytest_sliced.log
timeout -s KILL 420 python tools/batch_bench.py ${1:-100000000} ${2:-1024} 100 ${3:-} > gpurun_out/batch_bench.log 2>&1
echo "batch_bench rc=$?"; cat gpurun_out/batch_bench.log | tail -12
if [ "${4:-ncu}" = "ncu" ]; then
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:scan_sliced -s 1 -c 1 -f -o gpurun_out/prof_sliced python tools/prof_sliced.py 100000000 1024 2 > gpurun_out/ncu_sliced.log 2>&1
echo "ncu sliced rc=$?"; tail -2 gpurun_out/ncu_sliced.log
fi
