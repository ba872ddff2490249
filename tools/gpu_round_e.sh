#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_e.log
for nt in 64 128 256; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload traffic-64k --steps 50 --warmup 5 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_e_traffic-64k_nt$nt.json 2> gpurun_out/r02_bench_e_traffic-64k_nt$nt.err
done
for nt in 256 1024; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload large-1M --steps 10 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_e_large-1M_nt$nt.json 2> gpurun_out/r02_bench_e_large-1M_nt$nt.err
done
for nt in 128 256; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload train-py --steps 20 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_e_train-py_nt$nt.json 2> gpurun_out/r02_bench_e_train-py_nt$nt.err
done
python bench.py --steps 50 --warmup 5 --no-extra --cpu-seconds 0 --python-seconds 0 > gpurun_out/r02_bench_e_default.json 2> gpurun_out/r02_bench_e_default.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 4 --launch-count 1 -o gpurun_out/r02_large1M_tick_e -f \
  python bench.py --workload large-1M --steps 3 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_ncu_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 8 --launch-count 1 -o gpurun_out/r02_traffic64k_tick_e -f \
  python bench.py --workload traffic-64k --steps 5 --warmup 5 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_ncu_e2.log 2>&1
tail -5 gpurun_out/r02_pytest_e.log
