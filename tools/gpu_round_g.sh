#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_g_traffic-64k.json 2> gpurun_out/r02_bench_g_traffic-64k.err
python bench.py --workload train-py --steps 20 --warmup 3 $B > gpurun_out/r02_bench_g_train-py.json 2> gpurun_out/r02_bench_g_train-py.err
python bench.py --workload large-1M --steps 10 --warmup 3 $B > gpurun_out/r02_bench_g_large-1M.json 2> gpurun_out/r02_bench_g_large-1M.err
python bench.py --workload default-2M+final_observation --steps 50 --warmup 5 $B > gpurun_out/r02_bench_g_default_final.json 2> gpurun_out/r02_bench_g_default_final.err
tools/ncu_capture.sh r02g_large1M traffic_tick 4 --workload large-1M --steps 3 --warmup 3 $B
tools/ncu_capture.sh r02g_traffic64k traffic_tick 8 --workload traffic-64k --steps 5 --warmup 5 $B
tools/ncu_capture.sh r02g_trainpy traffic_tick 4 --workload train-py --steps 3 --warmup 3 $B
ls -la gpurun_out | tail -20
