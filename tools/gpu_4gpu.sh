set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 rc=$?"; cat gpurun_out/bench_n4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['config']['parallelism'], d['verified'], d.get('multi_query'))"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n4.err | tail -3
