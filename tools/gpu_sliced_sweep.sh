set -u
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:scan_sliced -s 1 -c 1 -f -o gpurun_out/prof_sliced python tools/prof_sliced.py 20000000 1024 2 > gpurun_out/ncu_sliced.log 2>&1
echo "ncu sliced rc=$?"; tail -2 gpurun_out/ncu_sliced.log
for cfg in "2000000 1024" "20000000 1024" "100000000 256" "100000000 64" "100000000 16"; do
  timeout -s KILL 300 python tools/batch_bench.py $cfg 100 fast 2>&1 | grep "bit-sliced"
done
