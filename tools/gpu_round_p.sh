#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
for mb in 16 12 10 8; do for cv in 100 50 35; do
PGTG_MAPGEN_MINB=$mb PGTG_MAPGEN_CARVEOUT=$cv python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_p_mb${mb}_cv${cv}.json 2> /dev/null
done; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_p.log
tail -3 gpurun_out/r02_pytest_p.log
