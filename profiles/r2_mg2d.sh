mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s -k "2gpu or oracle" > gpurun_out/r2_mg2d_tests.log 2>&1; grep -E "identical|passed|failed|rror" gpurun_out/r2_mg2d_tests.log | tail -8
run() { tag=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_$tag.json 2> gpurun_out/bench_r2e_$tag.err; echo "$tag rc=$?"; }
run dma X=1
run nccl OGB_ROWS_NCCL=1
python - <<'PY'
import json
for t in ('dma','nccl'):
    try:
        l=json.loads(open(f'gpurun_out/bench_r2e_{t}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(t, 'failed', e); continue
    print(t, round(l['ms_per_step'],2), l['parity'], {k:round(v,2) for k,v in l['phases_ms'].items()}, 'e2e', round(l['e2e']['ms_per_step'],1))
    print('   ', {k:round(v['ms_per_step'],2) for k,v in l['kernels'].items() if k.startswith('exch') or k in ('window_part','probe_parts','verify','hash_insert')})
PY
