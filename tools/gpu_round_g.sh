#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_g.log
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python bench.py --workload traffic-64k --steps 50 --warmup 5 $B > gpurun_out/r02_bench_g_traffic-64k.json 2> gpurun_out/r02_bench_g_traffic-64k.err
python bench.py --workload train-py --steps 20 --warmup 3 $B > gpurun_out/r02_bench_g_train-py.json 2> gpurun_out/r02_bench_g_train-py.err
python bench.py --workload large-1M --steps 10 --warmup 3 $B > gpurun_out/r02_bench_g_large-1M.json 2> gpurun_out/r02_bench_g_large-1M.err
python bench.py --workload default-2M+final_observation --steps 50 --warmup 5 $B > gpurun_out/r02_bench_g_default_final.json 2> gpurun_out/r02_bench_g_default_final.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 4 --launch-count 1 -o gpurun_out/r02_large1M_tick_g -f \
  python bench.py --workload large-1M --steps 3 --warmup 3 $B > gpurun_out/r02_ncu_g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 8 --launch-count 1 -o gpurun_out/r02_traffic64k_tick_g -f \
  python bench.py --workload traffic-64k --steps 5 --warmup 5 $B > gpurun_out/r02_ncu_g2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traffic_tick --launch-skip 4 --launch-count 1 -o gpurun_out/r02_trainpy_tick_g -f \
  python bench.py --workload train-py --steps 3 --warmup 3 $B > gpurun_out/r02_ncu_g3.log 2>&1
tail -5 gpurun_out/r02_pytest_g.log
