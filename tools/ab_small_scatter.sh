# A/B: single-launch small-network scatter (BE_B200_SMALL_SCATTER=0 -> compaction kernel + scatter kernel)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for S in 0 1; do
  BE_B200_SMALL_SCATTER=$S python bench.py --workload coba --steps 5 --warmup 3 --timesteps 100 --no-extra --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('small=$S coba us/timestep', round(d['ms_per_step']*1e3/100,2), 'kernel_us', round(d['roofline']['kernel_ms_avg']*1e3,2))
"
done
