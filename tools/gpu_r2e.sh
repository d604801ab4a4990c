#!/bin/bash
# Round 2, GPU pass (1 GPU): whole suite, ingest bench, ncu captures for profiles/.
set -u
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python tools/ingest_bench.py 50000000 4000000 > gpurun_out/ingest.log 2>&1
echo "ingest rc=$?"; cat gpurun_out/ingest.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -f -o gpurun_out/r02_prof_scan_200m python tools/prof_driver.py 200000000 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scan_topk -s 2 -c 1 --csv --log-file gpurun_out/r02_ncu_dram_1b.csv python tools/prof_driver.py 1000000000 4 > gpurun_out/ncu_dram_1b.log 2>&1
echo "ncu dram 1b rc=$?"; tail -3 gpurun_out/r02_ncu_dram_1b.csv
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_sliced -s 1 -c 1 -f -o gpurun_out/r02_prof_sliced_100m python tools/prof_sliced.py 100000000 1024 2 > gpurun_out/ncu_sliced.log 2>&1
echo "ncu sliced rc=$?"; tail -2 gpurun_out/ncu_sliced.log
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-full-verify > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; grep -c "scan_topk\|sliced\|merge" gpurun_out/r02_launches_bench_1b.csv
