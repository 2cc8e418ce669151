# A/B: index quads per lane of the 4-lane gather groups (BE_B200_GATHER_UQ = 5..8; default = ceil(row / 16))
python -m pytest tests -m gpu -x -q -k "csr or fcn or auto" 2>&1 | tail -2
for UQ in 8 7 6; do for WLD in csr_gather csr_gather_hetero; do
  BE_B200_GATHER_UQ=$UQ python bench.py --workload $WLD --steps 5 --warmup 3 --timesteps 50 --no-extra --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('UQ=$UQ $WLD', 'us/timestep', round(d['ms_per_step']*1e3/50,2), 'kernel_us', round(d['roofline']['kernel_ms_avg']*1e3,2), 'frac', d['roofline']['frac'])
"
done; done
