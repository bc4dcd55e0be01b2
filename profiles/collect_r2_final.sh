# Round-2 final evidence on one B200: full GPU test suite, bench lines, ncu launch list of the bench command, full ncu captures of
# k_window_part (with the summary persisting in L2) and of the simplification kernels. Run on the GPU box: bash profiles/collect_r2_final.sh
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -9 gpurun_out/final_gpu_tests.log
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
for cfg in 2 5; do timeout 200 python bench.py --config $cfg --steps 3 --no-cpu-baseline > gpurun_out/final_bench_config$cfg.json 2>> gpurun_out/final_bench.err; echo "config $cfg rc=$?"; done
CMD="timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2_final.csv $CMD > gpurun_out/final_ncu_launches.log 2>&1; echo "ncu list rc=$?"
CMD2="timeout 200 python profiles/exp.py --config 3 --steps 1 --warmup 0 --simplify"
$CMD2 > gpurun_out/final_plain2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^(k_window_part|k_c_ready|k_c_turns|k_c_jump)' -s 12 -c 10 -o gpurun_out/prof_r2_final $CMD2 > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/final_bench.json','gpurun_out/final_bench_config2.json','gpurun_out/final_bench_config5.json'):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(l['ms_per_step'],3), 'ms', round(l['value']/1e6,1), 'M reads/s parity', l['parity'], 'e2e', round(l['e2e']['ms_per_step'],2), 'roofline', l['roofline']['kernel'], round(l['roofline']['frac'],3), 'gather', round(l['roofline'].get('frac_random_gather') or 0,3), 'step frac', round(l['roofline_step']['frac'],3), 'simplify', (l.get('simplify') or {}).get('ms'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
