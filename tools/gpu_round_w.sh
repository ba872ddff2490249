#!/bin/bash
# Round-2 final single-GPU evidence: tests, the bench lines (both arms), the launch list, one full ncu capture per kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_default-2M.json 2> gpurun_out/r02_bench_default-2M.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_default-2M_steps20.json 2> gpurun_out/r02_bench_default-2M_steps20.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
for w in traffic-64k large-1M train-py; do
  python bench.py --workload $w --steps 30 --warmup 5 --cpu-seconds 3 --python-seconds 3 --e2e-steps 0 --no-extra > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 $B > gpurun_out/r02_launches.log 2>&1
tools/ncu_capture.sh r02_ncu_lean_tick "pgtg_tick_kernel" 6 --workload default-2M --steps 3 --warmup 3 $B
tools/ncu_capture.sh r02_ncu_mapgen mapgen_registers 6 --workload default-2M --steps 3 --warmup 3 $B
tools/ncu_capture.sh r02_ncu_traffic64k traffic_tick 8 --workload traffic-64k --steps 5 --warmup 5 $B
tools/ncu_capture.sh r02_ncu_large1M traffic_tick 4 --workload large-1M --steps 3 --warmup 3 $B
tools/ncu_capture.sh r02_ncu_trainpy traffic_tick 4 --workload train-py --steps 3 --warmup 3 $B
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > gpurun_out/r02_gpu.txt
tail -2 gpurun_out/r02_pytest_gpu.log
tools/ncu_capture.sh r02_ncu_lean_slide_tick pgtg_tick_kernel 6 --workload sliding-nsd-1M --steps 3 --warmup 3 $B
python bench.py --workload sliding-nsd-1M --steps 30 --warmup 5 --cpu-seconds 3 --python-seconds 3 --e2e-steps 0 --no-extra > gpurun_out/r02_bench_sliding-nsd-1M.json 2> gpurun_out/r02_bench_sliding-nsd-1M.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
tail -2 gpurun_out/r02_smoke.log
