#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_l.log
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 $B > gpurun_out/r02_bench_l_sliding_traffic_tick.json 2> gpurun_out/r02_bench_l_sliding_traffic_tick.err
PGTG_NO_TRAFFIC_KERNEL_CARFREE=1 python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 $B > gpurun_out/r02_bench_l_sliding_general.json 2> gpurun_out/r02_bench_l_sliding_general.err
PGTG_TRAFFIC_NT=256 python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 $B > gpurun_out/r02_bench_l_sliding_traffic_tick_nt256.json 2> /dev/null
tail -3 gpurun_out/r02_pytest_l.log
