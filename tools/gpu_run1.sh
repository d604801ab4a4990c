#!/bin/bash
# First GPU pass: smoke, parity tests, microbenchmarks, bench at 100 M and 1 B rows.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout -s KILL 300 tools/gsb_ubench 16 > gpurun_out/ubench.log 2>&1; cat gpurun_out/ubench.log
timeout -s KILL 600 python bench.py --rows 100000000 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_100m.json 2> gpurun_out/bench_100m.err
echo "bench100m rc=$?"; cat gpurun_out/bench_100m.json; tail -5 gpurun_out/bench_100m.err
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1b.json 2> gpurun_out/bench_1b.err
echo "bench1b rc=$?"; cat gpurun_out/bench_1b.json; tail -5 gpurun_out/bench_1b.err
