#!/bin/bash
# Round 2, third GPU pass (2 GPUs): multi-device tests, bench at N=2 (fused + NCCL), PDL A/B rerun.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -k "multi_device or recovers or reference_test_suite or reference_daemon" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
timeout -s KILL 300 python tools/pdl_ab.py 10000000 125000000 2>&1 | grep -v "GSB_PDL=0\|stable_query=False" | tee gpurun_out/pdl_ab_static.log
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; cat gpurun_out/bench_n2.json; tail -12 gpurun_out/bench_n2.err
GSB_FUSED_EXCHANGE=0 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-multi-query --no-single-process --no-full-verify > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
echo "bench n2 nccl rc=$?"; cat gpurun_out/bench_n2_nccl.json; tail -5 gpurun_out/bench_n2_nccl.err
