#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -f -o gpurun_out/prof_scan_200m python tools/prof_driver.py 200000000 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scan_topk -s 2 -c 1 --csv --log-file gpurun_out/ncu_dram_1b.csv python tools/prof_driver.py 1000000000 4 > gpurun_out/ncu_dram_1b.log 2>&1
echo "ncu dram 1b rc=$?"; tail -3 gpurun_out/ncu_dram_1b.csv
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout -s KILL 600 python tools/sweep.py 200000000 > gpurun_out/sweep.log 2>&1; grep -c same=True gpurun_out/sweep.log
timeout -s KILL 600 python tools/latency.py > gpurun_out/latency.log 2>&1; tail -22 gpurun_out/latency.log
timeout -s KILL 300 python tools/batch_bench.py 100000000 256 100 > gpurun_out/batch_bench.log 2>&1; cat gpurun_out/batch_bench.log
