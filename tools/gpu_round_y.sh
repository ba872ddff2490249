#!/bin/bash
mkdir -p gpurun_out
B="--cpu-seconds 0 --python-seconds 0 --e2e-steps 0 --no-extra"
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 $B > gpurun_out/r02_bench_y_$name.json 2> /dev/null; }
run base
run cv67_both PGTG_TICK_CARVEOUT=67 PGTG_MAPGEN_CARVEOUT=67
for g in 2 3 4 6; do
run cv67_grid$g PGTG_TICK_CARVEOUT=67 PGTG_MAPGEN_CARVEOUT=67 PGTG_MAPGEN_CTAS_PER_SM=$g
done
run cv75_grid3 PGTG_TICK_CARVEOUT=75 PGTG_MAPGEN_CARVEOUT=75 PGTG_MAPGEN_CTAS_PER_SM=3
run cv100_grid3 PGTG_MAPGEN_CARVEOUT=100 PGTG_MAPGEN_CTAS_PER_SM=3
run cv25_grid3 PGTG_MAPGEN_CTAS_PER_SM=3
