mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row['Kernel Name'][:44],[]).append(float(row['Metric Value'].replace(',','')))
tot=sum(sum(v)/len(v) for v in agg.values())
for k,v in agg.items(): print(f"{k:46s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us  share={sum(v)/len(v)/tot*100:5.1f}%")
PY
