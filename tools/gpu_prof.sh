#!/bin/bash
set -u
mkdir -p gpurun_out
export GSB_UNROLL=${GSB_UNROLL:-1}
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -f -o gpurun_out/prof_scan python tools/prof_driver.py 200000000 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --rows 100000000 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; tail -3 gpurun_out/ncu_bench.log; grep -c scan_topk gpurun_out/launches.csv
ls -la gpurun_out
