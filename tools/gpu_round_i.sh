#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_i.log
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload default-2M --steps 100 --warmup 5 $B > gpurun_out/r02_bench_i_default.json 2> gpurun_out/r02_bench_i_default.err
python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_i_traffic-64k.json 2> gpurun_out/r02_bench_i_traffic-64k.err
python bench.py --workload train-py --steps 20 --warmup 3 $B > gpurun_out/r02_bench_i_train-py.json 2> gpurun_out/r02_bench_i_train-py.err
python bench.py --workload large-1M --steps 10 --warmup 3 $B > gpurun_out/r02_bench_i_large-1M.json 2> gpurun_out/r02_bench_i_large-1M.err
tools/ncu_capture.sh r02i_mapgen mapgen 6 --workload default-2M --steps 3 --warmup 3 $B
tail -3 gpurun_out/r02_pytest_i.log
