#!/bin/bash
# Round 2, first GPU pass: smoke, the whole GPU suite, the PDL A/B, bench N=1.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout -s KILL 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python tools/pdl_ab.py > gpurun_out/pdl_ab.log 2>&1
echo "pdl_ab rc=$?"; cat gpurun_out/pdl_ab.log
timeout -s KILL 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cat gpurun_out/bench_n1.json; tail -15 gpurun_out/bench_n1.err
