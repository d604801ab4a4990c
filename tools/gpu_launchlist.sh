#!/bin/bash
# ncu launch list of the bench command (per-launch times are serialised and cold-cache: shares, not absolutes).
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; grep -c "gpu__time_duration" gpurun_out/launches_bench_1b.csv; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
