mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','parity','parity_detail','e2e','roofline','roofline_step','clocks','cpu_baseline','gather_ceilings'): print(k, json.dumps(l.get(k))[:700])
for k,v in l['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
