# A/B of the gather matvec's deferred hit weights (fp32 hetero): BE_B200_GATHER_DEFER_W=0 loads the weight quads in place
python -m pytest tests -m gpu -x -q -k "csr or fcn or auto" 2>&1 | tail -3
for R in off on; do for WLD in csr_gather csr_gather_hetero; do
  if [ $R = off ]; then export BE_B200_GATHER_DEFER_W=0; else unset BE_B200_GATHER_DEFER_W; fi
  python bench.py --workload $WLD --steps 5 --warmup 3 --timesteps 50 --no-extra --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('defer=$R $WLD', 'us/timestep', round(d['ms_per_step']*1e3/50,2), 'kernel_us', round(d['roofline']['kernel_ms_avg']*1e3,2), 'frac', d['roofline']['frac'], d['roofline']['kernel'])
"
done; done
