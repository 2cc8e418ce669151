#!/bin/bash
# One gpurun call (1 GPU): for every bench line, the plain run first (must exit 0), then one `ncu --set full` capture of 3
# launches of its dominant kernel.  Multi-GPU lines are captured on one GPU with the peers' exchange buffers emulated
# (--emulate-world N).  Summarise afterwards with tools/summarize_ncu.py <tag> <traffic key>.
# Usage: tools/ncu_all.sh [tag-prefix]      (default r2)
set -u
P=${1:-r2}
mkdir -p gpurun_out
COMMON="--steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph --no-e2e"
cap() {   # tag kernel-regex skip bench-args...
  local TAG=$1 KRE=$2 SKIP=$3; shift 3
  local CMD="python bench.py $COMMON $*"
  if $CMD > gpurun_out/${P}_${TAG}_plain.log 2>&1; then
    ncu --set full --clock-control none -k regex:${KRE} -s ${SKIP} -c 3 -o gpurun_out/${P}_${TAG}_prof -f $CMD > gpurun_out/${P}_${TAG}_ncu.log 2>&1
    # gpurun brings back at most 64 MiB: keep the raw metric table (text), drop the report
    ncu -i gpurun_out/${P}_${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_${TAG}_raw.csv 2>/dev/null
    rm -f gpurun_out/${P}_${TAG}_prof.ncu-rep gpurun_out/${P}_${TAG}_plain.log
    echo "$TAG: $(tail -1 gpurun_out/${P}_${TAG}_ncu.log | cut -c1-100) raw=$(wc -c < gpurun_out/${P}_${TAG}_raw.csv)"
  else
    echo "$TAG: plain run failed"; tail -3 gpurun_out/${P}_${TAG}_plain.log
  fi
}
cap fcn scatter_mv_kernel 8 --workload fcn
cap fcn_full scatter_mv_kernel 8 --workload fcn_full
cap fcn_w2 exchange_scatter 8 --workload fcn --emulate-world 2
cap fcn_w4 exchange_scatter 8 --workload fcn --emulate-world 4
cap fcn_w8 exchange_scatter 8 --workload fcn --emulate-world 8
cap fcn_full_w2 exchange_scatter 8 --workload fcn_full --emulate-world 2
cap fcn_full_w4 exchange_scatter 8 --workload fcn_full --emulate-world 4
cap coba scatter_mv_kernel 16 --workload coba
cap cuba_scatter jit_mv_scatter 16 --workload cuba_jitc
cap cuba_gather jit_mv_gather 8 --workload cuba_jitc_gather
cap jitc_scatter jit_mv_scatter 8 --workload jitc_1m
cap jitc_gather jit_mv_gather 4 --workload jitc_1m_gather
cap csr_scatter scatter_mv_kernel 8 --workload csr_scatter
cap csr_gather_homo gather_mv_kernel 8 --workload csr_gather
cap csr_gather_hetero gather_mv_kernel 8 --workload csr_gather_hetero
cap spmm scatter_mm_bm 12 --workload csr_spmm
cap dense dense_mm_t_tc 4 --workload dense_mm
for N in 2 4 8; do
  cap jitc_scatter_w$N jit_mv_scatter 8 --workload jitc_1m_sharded --emulate-world $N
  cap jitc_gather_w$N jit_mv_gather 4 --workload jitc_1m_sharded_gather --emulate-world $N
done
ls gpurun_out/${P}_*_raw.csv | wc -l; du -sh gpurun_out
