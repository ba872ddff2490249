#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_j.log
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload default-2M --steps 100 --warmup 5 $B > gpurun_out/r02_bench_j_default.json 2> gpurun_out/r02_bench_j_default.err
python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_j_traffic-64k.json 2> gpurun_out/r02_bench_j_traffic-64k.err
tail -3 gpurun_out/r02_pytest_j.log
