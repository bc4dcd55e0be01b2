timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { python bench.py --steps 5 --warmup 2 --no-cpu-baseline $* 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'phases',{k:round(v,3) for k,v in d['phases_ms'].items()},'probe_us',round(1e3*d['roofline']['kernel_ms'],1),'cands',d['stats']['candidates'])
    else: print(l.strip()[:300])
"; }
run
