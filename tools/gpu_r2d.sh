#!/bin/bash
# Round 2, GPU pass 4 (1 GPU): parity suite after the threshold-sharing change, A/B of GSB_SHARE_HIST.
set -u
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for sh in 0 1; do
  echo "== GSB_SHARE_HIST=$sh"
  GSB_SHARE_HIST=$sh timeout -s KILL 600 python tools/pdl_ab.py 10000000 125000000 1000000000 2>&1 | grep -v "GSB_PDL=0\|stable_query=False" | tee gpurun_out/pdl_ab_share$sh.log
done
