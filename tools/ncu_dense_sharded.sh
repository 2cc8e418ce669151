python -m pytest tests -m gpu -x -q -k "dense or sharded or c_host" 2>&1 | tail -2
for N in 2 8; do python bench.py --workload dense_sharded --emulate-world $N --steps 5 --warmup 3 --timesteps 50 --no-extra --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('dense_sharded world=$N', 'us/timestep', round(d['ms_per_step']*1e3/50,2), 'kernel_us', round(d['roofline']['kernel_ms_avg']*1e3,2), 'frac', d['roofline']['frac'])
"; done
P=r2b
COMMON="--steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph --no-e2e"
for N in 2 4 8; do
  TAG=dense_sharded_w$N
  CMD="python bench.py $COMMON --workload dense_sharded --emulate-world $N"
  ncu --set full --clock-control none -k regex:dense_mv_t_kernel -s 8 -c 3 -o gpurun_out/${P}_${TAG}_prof -f $CMD > gpurun_out/${P}_${TAG}_ncu.log 2>&1
  ncu -i gpurun_out/${P}_${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/${P}_${TAG}_prof.ncu-rep; echo "$TAG raw=$(wc -c < gpurun_out/${P}_${TAG}_raw.csv)"
done
