#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
python tools/scratch/inline_check.py > gpurun_out/r02_inline_check.log 2>&1
tail -2 gpurun_out/r02_inline_check.log
python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_aa_pipeline.json 2> /dev/null
PGTG_INLINE_MAPGEN=1 python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_aa_inline.json 2> gpurun_out/r02_bench_aa_inline.err
PGTG_INLINE_MAPGEN=1 PGTG_TICK_CARVEOUT=80 python bench.py --steps 40 --warmup 5 $B > gpurun_out/r02_bench_aa_inline_cv80.json 2> /dev/null
PGTG_INLINE_MAPGEN=1 python bench.py --workload default-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_aa_inline_64k.json 2> /dev/null
python bench.py --workload default-64k --steps 40 --warmup 5 $B > gpurun_out/r02_bench_aa_pipeline_64k.json 2> /dev/null
