#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_h.log
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_h_traffic-64k.json 2> gpurun_out/r02_bench_h_traffic-64k.err
PGTG_TRAFFIC_NT=256 python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_h_traffic-64k_nt256.json 2> /dev/null
python bench.py --workload train-py --steps 20 --warmup 3 $B > gpurun_out/r02_bench_h_train-py.json 2> gpurun_out/r02_bench_h_train-py.err
PGTG_TRAFFIC_NT=128 python bench.py --workload train-py --steps 20 --warmup 3 $B > gpurun_out/r02_bench_h_train-py_nt128.json 2> /dev/null
python bench.py --workload large-1M --steps 10 --warmup 3 $B > gpurun_out/r02_bench_h_large-1M.json 2> gpurun_out/r02_bench_h_large-1M.err
tail -3 gpurun_out/r02_pytest_h.log
