mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_r2c_8gpu.json 2> gpurun_out/bench_r2c_8gpu.err; echo "bench8 rc=$?"; tail -5 gpurun_out/bench_r2c_8gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2c_8gpu.json',):
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f)
    for k in ('value','ms_per_step','parity','phases_ms','setup_s'): print(' ',k, json.dumps(l.get(k))[:600])
    print('  e2e', l['e2e']['ms_per_step'], l['e2e']['value'])
    print('  ', {k:round(v['ms_per_step'],3) for k,v in l['kernels'].items()})
PY
