#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_c.log
for nt in 64 128 256; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload traffic-64k --steps 50 --warmup 5 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_c_traffic-64k_nt$nt.json 2> gpurun_out/r02_bench_c_traffic-64k_nt$nt.err
done
for nt in 256 1024; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload large-1M --steps 10 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_c_large-1M_nt$nt.json 2> gpurun_out/r02_bench_c_large-1M_nt$nt.err
done
for nt in 128 256; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload train-py --steps 20 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_c_train-py_nt$nt.json 2> gpurun_out/r02_bench_c_train-py_nt$nt.err
done
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_c_full.json 2> gpurun_out/r02_bench_c_full.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_c_reference.json 2> gpurun_out/r02_bench_c_reference.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 4 --launch-count 1 -o gpurun_out/r02_large1M_tick -f \
  python bench.py --workload large-1M --steps 3 --warmup 3 --cpu-seconds 0 --python-seconds 0 --e2e-steps 0 > gpurun_out/r02_ncu_c.log 2>&1
tail -3 gpurun_out/r02_pytest_c.log
