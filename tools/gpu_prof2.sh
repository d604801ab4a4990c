#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -f -o gpurun_out/prof_scan_200m python tools/prof_driver.py 200000000 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scan_topk -s 2 -c 1 --csv --log-file gpurun_out/ncu_dram_1b.csv python tools/prof_driver.py 1000000000 4 > gpurun_out/ncu_dram_1b.log 2>&1
echo "ncu dram 1b rc=$?"; tail -4 gpurun_out/ncu_dram_1b.csv
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
python tools/sweep.py 10000000 > gpurun_out/sweep_10m.log 2>&1; grep "rowpop=1 warps=1[26] stages=2" gpurun_out/sweep_10m.log
python tools/sweep.py 10000000 10 > gpurun_out/sweep_10m_k10.log 2>&1; grep "rowpop=1 warps=1[26] stages=2" gpurun_out/sweep_10m_k10.log
