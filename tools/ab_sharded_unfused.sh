for N in 2 8; do for F in 1 0; do
export BE_B200_SHARDED_FUSED=$F
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scatter_mv|exchange_scatter|wait_copy|pack_publish|compact" -s 40 -c 40 --csv --log-file gpurun_out/r2c_unf_${N}_${F}.csv python bench.py --workload fcn --emulate-world $N --steps 2 --warmup 3 --timesteps 8 --no-extra --no-cpu --no-graph --no-e2e > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2c_unf_${N}_${F}.csv")) if len(r)>10]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki].split("(")[0][:60],[]).append(float(r[vi].replace(",","")))
    except ValueError: pass
print("world=$N fused=$F", {k:(len(v), round(sum(v)/len(v)/1e3,2)) for k,v in agg.items()})
PY
done; done
