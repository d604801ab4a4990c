#!/bin/bash
# 2-GPU pass: fused exchange.  GPU tests (incl. the single-GPU two-rank fused test), then bench N=2 fused vs NCCL.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for fused in 1 0; do
GSB_FUSED_EXCHANGE=$fused timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_fused$fused.json 2> gpurun_out/bench_n2_fused$fused.err
echo "bench n2 fused=$fused rc=$?"; cat gpurun_out/bench_n2_fused$fused.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['config']['parallelism'], d['verified'])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2_fused$fused.err | tail -5
done
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cat gpurun_out/bench_n1.json
