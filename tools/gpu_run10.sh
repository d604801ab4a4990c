#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/ingest_bench.py 20000000 2>&1 | tail -5
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --rows 125000000 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
