set -u
timeout -s KILL 600 python -m pytest tests -q -m gpu -k "batch_kernel_choice or sliced_kernel_padded or sharded_multi_query" --timeout 300 2>&1 | tail -2
timeout -s KILL 300 python tools/crossover.py 100000000 2>&1 | grep rows=
timeout -s KILL 300 python tools/crossover.py 10000000 2>&1 | grep rows=
