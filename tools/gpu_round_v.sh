#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
for w in traffic-64k train-py; do
python bench.py --workload $w --steps 40 --warmup 5 $B > gpurun_out/r02_bench_v_$w.json 2> /dev/null
PGTG_MAPGEN_MINB=16 python bench.py --workload $w --steps 40 --warmup 5 $B > gpurun_out/r02_bench_v_${w}_mb16.json 2> /dev/null
done
python bench.py --workload default-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_v_default-64k.json 2> /dev/null
PGTG_MAPGEN_CARVEOUT=100 python bench.py --workload default-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_v_default-64k_cv100.json 2> /dev/null
