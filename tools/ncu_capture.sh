#!/bin/bash
# One gpurun call: plain run first (must exit 0), then the ncu launch list and one full capture
# of the dominant kernel.  Usage: tools/ncu_capture.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; shift
KRE=$1; shift
CMD="python bench.py --steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 3 \
    -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "plain:"; tail -c 600 gpurun_out/${TAG}_plain.log; echo; tail -3 gpurun_out/${TAG}_ncu1.log; tail -3 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/
