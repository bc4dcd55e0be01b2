# Round 2, 8 x B200: the 8-rank parity tests (adversarial sets against the oracle; config 2 x 8, config 3 x 8 and config 4 at
# full size against the goldens), the 4-rank config 3 x 4 case, and the bench at 8 and 4 GPUs.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s -k "8gpu or oracle[8] or 4gpu-config3" > gpurun_out/r2_mg8b_tests.log 2>&1; grep -E "identical|passed|failed|rror|skipped" gpurun_out/r2_mg8b_tests.log | tail -30
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; echo "bench8 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29411 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r2_4gpu.json 2> gpurun_out/bench_r2_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2_8gpu.json','gpurun_out/bench_r2_4gpu.json'):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'failed', e); continue
    print(f)
    for k in ('value','ms_per_step','parity','phases_ms','setup_s'): print(' ',k, json.dumps(l.get(k))[:600])
    print('  e2e', l['e2e']['ms_per_step'], l['e2e']['value'])
    print('  ', {k:round(v['ms_per_step'],3) for k,v in l['kernels'].items()})
PY
