#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 600 python tools/sweep.py 200000000 > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1b.json 2> gpurun_out/bench_1b.err
echo "bench1b rc=$?"; cat gpurun_out/bench_1b.json; tail -5 gpurun_out/bench_1b.err
