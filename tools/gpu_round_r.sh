#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_r.log
python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_r_default.json 2> gpurun_out/r02_bench_r_default.err
for cv in 35 50; do PGTG_MAPGEN_CARVEOUT=$cv python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_r_cv$cv.json 2> /dev/null; done
PGTG_MAPGEN_MINB=16 python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_r_mb16.json 2> /dev/null
bash tools/ncu_capture.sh r02_mapgen_r pgtg_mapgen_registers_kernel 6 --steps 3 --warmup 3 $B
tail -3 gpurun_out/r02_pytest_r.log
