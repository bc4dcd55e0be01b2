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
