#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <launch-skip> <bench args...>
# One `ncu --set full` capture of a kernel of bench.py; keeps the summaries and the gzipped source-page csv under gpurun_out/
# (the .ncu-rep itself is tens of MB: over the size limit of what a gpurun call brings back).
name=$1; regex=$2; skip=$3; shift 3
rep=/tmp/$name.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip --launch-count 1 -o /tmp/$name -f \
  python bench.py "$@" > gpurun_out/${name}_ncu.log 2>&1
ncu -i $rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/${name}_src.csv.gz
python tools/ncu_summary.py $rep 40 > gpurun_out/${name}_summary.txt 2>&1
