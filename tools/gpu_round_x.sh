#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_x.log
python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 $B > gpurun_out/r02_bench_x_sliding_lean.json 2> gpurun_out/r02_bench_x_sliding_lean.err
PGTG_NO_LEAN=1 python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 $B > gpurun_out/r02_bench_x_sliding_general.json 2> /dev/null
python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_x_default.json 2> /dev/null
bash tools/ncu_capture.sh r02_ncu_lean_slide_tick pgtg_tick_kernel 6 --workload sliding-nsd-1M --steps 3 --warmup 3 $B
tail -3 gpurun_out/r02_pytest_x.log
