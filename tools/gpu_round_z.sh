#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done >> gpurun_out/r02_topo.txt
nproc >> gpurun_out/r02_topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --cpu-seconds 0 --python-seconds 0 --no-extra > gpurun_out/r02_bench_${N}gpu_numa.json 2> gpurun_out/r02_bench_${N}gpu_numa.err
echo "rc=$?"
PGTG_NO_NUMA_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --cpu-seconds 0 --python-seconds 0 --no-extra > gpurun_out/r02_bench_${N}gpu_nobind.json 2> gpurun_out/r02_bench_${N}gpu_nobind.err
echo "rc=$?"
