#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_u.log
python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_u_default.json 2> /dev/null
PGTG_MAPGEN_MINB=8 python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_u_default_mb8.json 2> /dev/null
for w in traffic-64k train-py large-1M; do
python bench.py --workload $w --steps 30 --warmup 5 $B > gpurun_out/r02_bench_u_$w.json 2> /dev/null
PGTG_NO_MAP_IN_REGISTERS=1 python bench.py --workload $w --steps 30 --warmup 5 $B > gpurun_out/r02_bench_u_${w}_staged.json 2> /dev/null
done
tail -3 gpurun_out/r02_pytest_u.log
