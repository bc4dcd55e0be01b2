mkdir -p gpurun_out
for S in 0 1; do
CMD="python bench.py --scale 4 --steps 1 --warmup 1 --no-cpu-baseline"
OGB_SUMMARY=$S $CMD > gpurun_out/plain.log 2>&1 && OGB_SUMMARY=$S ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_big$S.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "== OGB_SUMMARY=$S (scale 4: 5.2M reads, 332 MB index)"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/plain.log | head -1
python - <<PY
import csv,collections
lines=[l for l in open('gpurun_out/launches_big$S.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    k=(row['Kernel Name'][:34],row['Metric Name'])
    agg.setdefault(k,[]).append(float(row['Metric Value'].replace(',','')))
names=collections.OrderedDict((k[0],1) for k in agg)
for n in names:
    t=agg[(n,'gpu__time_duration.sum')]; b=agg.get((n,'dram__bytes_read.sum'),[0])
    print(f"{n:36s} n={len(t):4d} mean={sum(t)/len(t)/1e3:9.1f} us  dram_read_mean={sum(b)/len(b)/1e6:9.1f} MB")
PY
done
