#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_s_overlap.json 2> /dev/null
PGTG_NO_OVERLAP=1 python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_s_serial.json 2> /dev/null
PGTG_NO_OVERLAP=1 PGTG_MAPGEN_CARVEOUT=100 python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_s_serial_cv100.json 2> /dev/null
python bench.py --workload traffic-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_s_traffic_overlap.json 2> /dev/null
PGTG_NO_OVERLAP=1 python bench.py --workload traffic-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_s_traffic_serial.json 2> /dev/null
