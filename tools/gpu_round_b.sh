#!/bin/bash
# GPU run: tests, traffic benches with CTA-size variants, one full ncu capture of the traffic tick
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_b.log
for nt in 128 256; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload traffic-64k --steps 50 --warmup 5 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_b_traffic-64k_nt$nt.json 2> gpurun_out/r02_bench_b_traffic-64k_nt$nt.err
done
for nt in 256 1024; do
  PGTG_TRAFFIC_NT=$nt python bench.py --workload large-1M --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_b_large-1M_nt$nt.json 2> gpurun_out/r02_bench_b_large-1M_nt$nt.err
done
python bench.py --workload default-2M --steps 30 --warmup 5 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/r02_bench_b_default-2M.json 2> gpurun_out/r02_bench_b_default-2M.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 8 --launch-count 1 -o gpurun_out/r02_traffic64k_tick -f \
  python bench.py --workload traffic-64k --steps 5 --warmup 5 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/r02_ncu_b.log 2>&1
tail -3 gpurun_out/r02_pytest_b.log
