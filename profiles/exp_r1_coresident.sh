run() { python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'k3',round(d['phases_ms']['ms_scan_kernel'],3),'probe_launch_us',round(1e3*d['roofline']['kernel_ms'],1))
    else: print(l.strip()[:300])
"; }
for cfg in "0 0" "4 4" "3 3" "4 2" "3 2" "2 2" "5 3" "6 2"; do set -- $cfg; echo "== probe blocks/SM=$1 verify blocks/SM=$2"; OGB_PROBE_BLOCKS_PER_SM=$1 OGB_VERIFY_BLOCKS_PER_SM=$2 run; done
