#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_o_default.json 2> gpurun_out/r02_bench_o_default.err
for cv in 0 50; do
PGTG_MAPGEN_CARVEOUT=$cv python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_o_carve$cv.json 2> /dev/null
done
PGTG_NO_MAP_IN_REGISTERS=1 python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_o_shared_maps.json 2> /dev/null
python bench.py --workload traffic-64k --steps 30 --warmup 5 $B > gpurun_out/r02_bench_o_traffic.json 2> /dev/null
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_o.log
bash tools/ncu_capture.sh r02_mapgen_o pgtg_mapgen_kernel 6 --steps 3 --warmup 3 $B
tail -3 gpurun_out/r02_pytest_o.log
