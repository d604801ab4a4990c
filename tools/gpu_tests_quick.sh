#!/bin/bash
# whole GPU suite + smoke + a 2-step bench line (round-end style check on one GPU)
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['e2e']['value'], d['verified'], d['multi_query']['batch_ms'], d['roofline']['frac'])"
