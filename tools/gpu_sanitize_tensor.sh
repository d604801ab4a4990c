#!/bin/bash
# compute-sanitizer over the tensor-core multi-query kernel (small cases) and the single-query kernel
# with the barrier-free select: memcheck, racecheck, synccheck.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  TC_CASE=2 timeout -s KILL 500 compute-sanitizer --tool $tool --kernel-regex kns=scan_tensor python tools/tensor_try.py > gpurun_out/sanitize_tensor_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|PARITY|identical" gpurun_out/sanitize_tensor_$tool.log | tail -4
done
GSB_TAIL=1 timeout -s KILL 400 compute-sanitizer --tool memcheck --kernel-regex kns=scan_topk python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ties or sizes_and_ragged" > gpurun_out/sanitize_tail_memcheck.log 2>&1
echo "tail memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_tail_memcheck.log | tail -3
GSB_BATCH_KERNEL=3 timeout -s KILL 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sliced_kernel_ties_blocks_and_groups or sharded_multi_query_merge" > gpurun_out/sanitize_sliced_memcheck.log 2>&1
echo "sliced memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_sliced_memcheck.log | tail -3
