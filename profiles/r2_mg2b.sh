mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -m gpu -s > gpurun_out/r2_mg2b_tests.log 2>&1; grep -E "identical|passed|failed|Error|error" gpurun_out/r2_mg2b_tests.log | tail -12
python bench.py --steps 4 --no-cpu-baseline > gpurun_out/bench_r2b_1gpu.json 2> gpurun_out/bench_r2b_1gpu.err; echo "bench1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_r2b_2gpu.json 2> gpurun_out/bench_r2b_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_r2b_2gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2b_1gpu.json','gpurun_out/bench_r2b_2gpu.json'):
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f)
    for k in ('value','ms_per_step','parity','phases_ms'): print(' ',k, json.dumps(l.get(k))[:600])
    print('  e2e', l['e2e']['ms_per_step'], l['e2e']['value'])
    print('  ', {k:round(v['ms_per_step'],3) for k,v in l['kernels'].items()})
PY
