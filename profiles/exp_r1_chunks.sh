timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'phases',{k:round(v,3) for k,v in d['phases_ms'].items()},'k3',round(d['roofline']['kernel_ms'],3))
    else: print(l.strip()[:300])
"; }
for C in 8192 16384 32768 65536; do echo "== CHUNK=$C two streams"; OGB_CHUNK_READS=$C run; done
echo "== CHUNK=65536 one stream"; OGB_ONE_STREAM=1 OGB_CHUNK_READS=65536 run
echo "== CHUNK=16384 one stream"; OGB_ONE_STREAM=1 OGB_CHUNK_READS=16384 run
