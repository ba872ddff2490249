#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
for mb in 10 8; do for cv in 20 25 30; do
PGTG_MAPGEN_MINB=$mb PGTG_MAPGEN_CARVEOUT=$cv python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_q_mb${mb}_cv${cv}.json 2> /dev/null
done; done
for w in traffic-64k default-2M+final_observation train-py; do
PGTG_MAPGEN_MINB=10 PGTG_MAPGEN_CARVEOUT=35 python bench.py --workload $w --steps 30 --warmup 5 $B > gpurun_out/r02_bench_q_${w}_mb10_cv35.json 2> /dev/null
done
