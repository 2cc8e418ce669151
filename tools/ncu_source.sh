#!/bin/bash
# Per-instruction stall sampling of one kernel: ncu --set full --import-source on, source page exported as CSV on the box.
# Usage: tools/ncu_source.sh <tag> <kernel-regex> <skip> <bench args...>
set -u
TAG=$1 KRE=$2 SKIP=$3; shift 3
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph --no-e2e $*"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:${KRE} -s ${SKIP} -c 1 -o gpurun_out/${TAG}_src -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_src.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>/dev/null
rm -f gpurun_out/${TAG}_src.ncu-rep gpurun_out/${TAG}_plain.log
wc -c gpurun_out/${TAG}_source_sass.csv
