#!/bin/bash
# Refresh of the captures whose kernels changed after tools/ncu_all.sh ran (tag prefix r2b), the dense post-sharded lines,
# and the launch list of the default bench command.  One GPU.
set -u
P=${1:-r2b}
mkdir -p gpurun_out
COMMON="--steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph --no-e2e"
cap() {   # tag kernel-regex skip bench-args...
  local TAG=$1 KRE=$2 SKIP=$3; shift 3
  local CMD="python bench.py $COMMON $*"
  if $CMD > gpurun_out/${P}_${TAG}_plain.log 2>&1; then
    ncu --set full --clock-control none -k regex:${KRE} -s ${SKIP} -c 3 -o gpurun_out/${P}_${TAG}_prof -f $CMD > gpurun_out/${P}_${TAG}_ncu.log 2>&1
    ncu -i gpurun_out/${P}_${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_${TAG}_raw.csv 2>/dev/null
    rm -f gpurun_out/${P}_${TAG}_prof.ncu-rep gpurun_out/${P}_${TAG}_plain.log
    echo "$TAG: raw=$(wc -c < gpurun_out/${P}_${TAG}_raw.csv)"
  else
    echo "$TAG: plain run failed"; tail -3 gpurun_out/${P}_${TAG}_plain.log
  fi
}
cap csr_gather_homo gather_mv_kernel 8 --workload csr_gather
cap csr_gather_hetero gather_mv_kernel 8 --workload csr_gather_hetero
cap dense dense_mm_t_tc 4 --workload dense_mm
for N in 2 4 8; do cap dense_sharded_w$N dense_mv_t_kernel 8 --workload dense_sharded --emulate-world $N; done
# launch list of the default bench command (steps reduced so that ncu's serialisation stays short): shares, not absolutes
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-e2e > gpurun_out/${P}_fcn_plain.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_fcn_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-e2e > /dev/null 2>&1
wc -l gpurun_out/${P}_fcn_launches.csv
# (later in round 2 the same `cap` was run for: coba -> scatter_mv_small, jitc_1m_gather / cuba_jitc_gather -> jit_mv_gather)
