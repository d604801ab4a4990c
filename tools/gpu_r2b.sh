#!/bin/bash
# Round 2, second GPU pass (1 GPU): whole GPU suite, candidate-buffer size A/B.
set -u
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
for cap in 4096 8192; do
  echo "== GSB_MIN_CAP=$cap"
  GSB_MIN_CAP=$cap timeout -s KILL 600 python tools/pdl_ab.py 10000000 125000000 1000000000 2>&1 | grep -v "GSB_PDL=0\|stable_query=False" | tee gpurun_out/pdl_ab_cap$cap.log
done
