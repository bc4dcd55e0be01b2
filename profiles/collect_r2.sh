# Round-2 evidence on one B200: full GPU test suite, bench lines (ours + reference arm), ncu launch list of the bench
# command, full ncu captures of the hot kernels. Run on the GPU box: bash profiles/collect_r2.sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/gpu_tests_r2.log 2>&1; tail -9 gpurun_out/gpu_tests_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2.err; echo "reference rc=$?"
CMD="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMD2="timeout 300 python profiles/exp.py --config 3 --steps 1 --warmup 0"
$CMD2 > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(k_window_part|k_probe_parts|k_verify|k_key_part|k_insert_parts|k_rows_finish)' -s 30 -c 8 -o gpurun_out/prof_r2_scan $CMD2 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(k_mark|k_keep|k_emit)' -s 6 -c 6 -o gpurun_out/prof_r2_reduce $CMD2 >> gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
head -c 600 gpurun_out/bench_r2.json
# the other single-GPU configurations through the same bench (parity field, mixed-length upload path, device Dataset stage for the record)
for cfg in 2 5; do timeout 300 python bench.py --config $cfg --steps 3 --no-cpu-baseline > gpurun_out/bench_r2_config$cfg.json 2>> gpurun_out/bench_r2.err; echo "config $cfg rc=$?"; done
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2.json','gpurun_out/bench_r2_config2.json','gpurun_out/bench_r2_config5.json'):
    l=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(l['ms_per_step'],3), 'ms', round(l['value']/1e6,1), 'M reads/s parity', l['parity'], 'e2e', round(l['e2e']['ms_per_step'],2), 'roofline', l['roofline']['kernel'], round(l['roofline']['frac'],3), 'gather', round(l['roofline'].get('frac_random_gather') or 0,3), 'step frac', round(l['roofline_step']['frac'],3), 'setup', l['setup_s'])
PY
