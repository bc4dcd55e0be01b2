mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_$tag.json 2> gpurun_out/bench_r2d_$tag.err; echo "$tag rc=$?"; }
run base X=1
run ce NCCL_P2P_USE_CUDA_MEMCPY=1
run ctas4 NCCL_MAX_CTAS=4
run ctas2 NCCL_MAX_CTAS=2
python - <<'PY'
import json
for t in ('base','ce','ctas4','ctas2'):
    try:
        l=json.loads(open(f'gpurun_out/bench_r2d_{t}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(t, 'failed', e); continue
    print(t, round(l['ms_per_step'],2), l['parity'], {k:round(v,2) for k,v in l['phases_ms'].items()}, 'e2e', round(l['e2e']['ms_per_step'],1))
    print('   ', {k:round(v['ms_per_step'],2) for k,v in l['kernels'].items() if k.startswith('exch') or k in ('window_part','probe_parts','verify')})
PY
