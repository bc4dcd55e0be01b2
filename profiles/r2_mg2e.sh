mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s -k "2gpu or oracle" > gpurun_out/r2_mg2e_tests.log 2>&1; grep -E "passed|failed|rror" gpurun_out/r2_mg2e_tests.log | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2_2gpu.json').read().strip().splitlines()[-1])
print(round(l['ms_per_step'],2), l['parity'], {k:round(v,2) for k,v in l['phases_ms'].items()}, 'e2e', round(l['e2e']['ms_per_step'],1))
print('   ', {k:round(v['ms_per_step'],2) for k,v in l['kernels'].items()})
PY
