timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for T in 1 2 4; do
echo "== OGB_TILE_READS=$T"
OGB_TILE_READS=$T python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'phases',{k:round(v,3) for k,v in d['phases_ms'].items()},'k3',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'sectors/probe',round(d['stats']['probe_sectors']/d['stats']['overlap_probes'],3), 'overflow', d['stats']['overflow_reads'])
    else: print(l.strip()[:300])
"
done
