mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s > gpurun_out/r2_mg2_tests.log 2>&1; tail -12 gpurun_out/r2_mg2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_2gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2_2gpu.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','parity','e2e','roofline_step','phases_ms'): print(k, json.dumps(l.get(k))[:600])
for k,v in l['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','launches_per_step')})
PY
