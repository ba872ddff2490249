#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_f.log
B="--steps 50 --warmup 5 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
for cfg in "32 64" "32 128" "64 64" "64 128" "128 128" "128 256"; do set -- $cfg
  PGTG_TRAFFIC_G=$1 PGTG_TRAFFIC_NT=$2 python bench.py --workload traffic-64k $B > gpurun_out/r02_bench_f_traffic-64k_g$1_nt$2.json 2> gpurun_out/r02_bench_f_traffic-64k_g$1_nt$2.err
done
for cfg in "32 128" "32 256" "64 128" "64 256" "128 256"; do set -- $cfg
  PGTG_TRAFFIC_G=$1 PGTG_TRAFFIC_NT=$2 python bench.py --workload train-py --steps 20 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_f_train-py_g$1_nt$2.json 2> gpurun_out/r02_bench_f_train-py_g$1_nt$2.err
done
PGTG_TRAFFIC_NT=1024 python bench.py --workload large-1M --steps 10 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_f_large-1M_nt1024.json 2> gpurun_out/r02_bench_f_large-1M_nt1024.err
python bench.py --workload default-2M+final_observation $B > gpurun_out/r02_bench_f_default_final.json 2> gpurun_out/r02_bench_f_default_final.err
tail -5 gpurun_out/r02_pytest_f.log
