#!/bin/bash
# multi-GPU check of the final tree: the driver's own launch line at N = 2 (and whatever --gpus gave us)
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "rc=$?"
tail -c 600 gpurun_out/r02_bench_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_bench_${N}gpu_reference.json 2> gpurun_out/r02_bench_${N}gpu_reference.err
echo "rc=$?"
